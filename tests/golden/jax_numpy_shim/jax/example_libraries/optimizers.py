"""name only: vqmc.py imports it at module level"""


def adam(*a, **k):
    raise NotImplementedError
