"""names only: the bijections that use jax.scipy.linalg are outside the spline hot path"""
