#!/bin/bash
# usage: tools/sass_extract.sh > profiles/r2_sass_extract.txt
# Counts, per kernel of the built lib/libmf.so, the SASS mnemonics that prove what the kernels use (B200_PROFILING.md):
# UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTMALDG (TMA tensor loads), FFMA2/FMUL2 (packed fp32), LDGSTS (cp.async),
# SYNCS (mbarrier), MEMBAR (fences).
LIB="$(cd "$(dirname "$0")/.." && pwd)/question-recommendation-system_b200/lib/libmf.so"
echo "# cuobjdump -sass $LIB  (sha256 $(sha256sum "$LIB" | cut -c1-16), $(date -u +%F))"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; next }
  { for (i = 1; i <= NF; i++) { t = $i; sub(/\..*/, "", t);
      if (t ~ /^(UTCHMMA|LDTM|UTMALDG|UTMAPF|FFMA2|FMUL2|LDGSTS|SYNCS|MEMBAR|ATOMS|UBLKPF)$/) c[fn " " t]++ } }
  END { for (k in c) print k, c[k] }' | sort | c++filt | awk '{ n=$NF; op=$(NF-1); $NF=""; $(NF-1)=""; printf "%-8s %5d  %s\n", op, n, $0 }' | cut -c1-200
