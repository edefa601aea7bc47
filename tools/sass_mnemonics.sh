#!/bin/bash
# Counts of the Blackwell-native SASS mnemonics per object of the shipped library -> profiles/sass_mnemonics.txt
cd "$(dirname "$0")/.."
out=profiles/sass_mnemonics.txt
{
echo "# cuobjdump -sass waveflow_b200/csrc/_obj/*.o | grep -c <mnemonic>   (objects of libwaveflow_b200.so, sm_100a)"
echo "# UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk,"
echo "# LDGSTS = cp.async, HMMA would be the legacy mma.sync path (none), SYNCS = mbarrier"
printf "%-34s %8s %6s %6s %7s %8s %7s %7s %6s %6s\n" object UTCHMMA LDTM STTM UTCBAR UTMALDG UBLKCP LDGSTS HMMA SYNCS
for f in waveflow_b200/csrc/_obj/*.o; do
  s=$(cuobjdump -sass "$f" 2>/dev/null)
  c() { echo "$s" | grep -c "$1"; }
  printf "%-34s %8s %6s %6s %7s %8s %7s %7s %6s %6s\n" "$(basename $f)" "$(c 'UTC[A-Z]*MMA')" "$(c LDTM)" "$(c STTM)" "$(c UTCBAR)" "$(c UTMALDG)" "$(c UBLKCP)" "$(c LDGSTS)" "$(c ' HMMA')" "$(c SYNCS)"
done
} > $out
cat $out
