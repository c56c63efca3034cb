#!/bin/bash
# SASS digest of libsong_los.so: how often the opcodes that prove the design (FP64 tensor MMA, TMA / bulk copies,
# mbarrier) occur per kernel.  usage: tools/sass_digest.sh > profiles/rNN_sass_digest.txt
SO=${1:-song_b200/libsong_los.so}
echo "# cuobjdump -sass $SO (sm_100a), opcode counts per kernel"
cuobjdump -sass "$SO" | awk '
  /Function :/ { fn=$3; next }
  { for (i = 1; i <= NF; i++) if ($i ~ /^(DMMA|DFMA|DMUL|DADD|UTMALDG|UBLKCP|SYNCS|UTCMMA|LDTM|HMMA|LDGSTS|REDUX|SHFL|BAR)/) { split($i, a, "."); key = fn " " $i; n[key]++ } }
  END { for (k in n) print n[k], k }' | sort -k2,2 -k1,1nr | awk '{ printf "%-90s %-28s %6d\n", $2, $3, $1 }' | c++filt | grep -E "los_tc_kernel<12, 1, 7|los_tc_kernel<4, 2, 8|los_exact_kernel|tc_interp|proj_dense" | head -120
echo
echo "# totals over the library"
cuobjdump -sass "$SO" | grep -oE "\b(DMMA\.[0-9x]+|UTMALDG\.[0-9A-Z.]+|UBLKCP\.[A-Z.]+|SYNCS\.[A-Z_.0-9]+|UTCMMA[.A-Z0-9]*|LDTM[.A-Z0-9]*)" | sort | uniq -c | sort -nr
echo
echo "# TMA / bulk copies that carry an L2 cache-policy descriptor (createpolicy evict_first / evict_last, los_tc.cu)"
cuobjdump -sass "$SO" | grep -E "UTMALDG|UBLKCP" | awk '{ t++; if ($0 ~ /desc\[/) d++ } END { printf "%d of %d\n", d, t }'
