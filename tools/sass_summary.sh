#!/bin/bash
# Per-kernel counts of the SASS opcodes that prove the Blackwell-native path (B200_PROFILING.md: tcgen05.mma ->
# UTC*MMA, tcgen05.ld -> LDTM, TMA -> UTMALDG/UTMASTG/UTMAREDG/UBLKCP).  Usage: tools/sass_summary.sh > profiles/sass_rNN.txt
set -e
ROOT="$(cd "$(dirname "${BASH_SOURCE[0]}")/.." && pwd)"
SO="$ROOT/instance-segment-basi_b200/basi_b200/libbasi_b200.so"
echo "# cuobjdump -sass $(basename "$SO") ($(date -u +%Y-%m-%dT%H:%MZ)); counts per kernel of the Blackwell-only opcodes"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3 }
  { for (i = 1; i <= NF; ++i) if ($i ~ /^(UTCHMMA|UTCQMMA|UTCBAR|UTCATOMSWS|LDTM|STTM|UTMALDG|UTMASTG|UTMAREDG|UBLKCP|UTMAPF|HMMA|SYNCS)/) { op=$i; sub(/;$/, "", op); c[fn"\t"op]++; t[op]++ } }
  END { for (k in c) print c[k] "\t" k; print "---- totals"; for (o in t) print t[o] "\t" o }' | sort -t$'\t' -k2,2 -k3,3 | c++filt | cut -c1-220
