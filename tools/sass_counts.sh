#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the tcgen05 / TMEM / TMA path (B200_PROFILING.md):
#   UTCHMMA = tcgen05.mma (bf16), LDTM = tcgen05.ld, UTMALDG / UTMASTG = TMA tensor load / store,
#   UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.   usage: tools/sass_counts.sh > profiles/rN_sass_counts.txt
LIB=${1:-3dgan_b200/lib/libb200gan.so}
echo "# cuobjdump -sass $LIB | per-kernel mnemonic counts (sm_100a)"
cuobjdump -sass "$LIB" | awk '
  /Function :/ { fn=$3; next }
  { for (m in pat) if ($0 ~ pat[m]) c[fn, m]++ }
  BEGIN { pat["UTCHMMA.2CTA"]="UTCHMMA\\.2CTA"; pat["UTCHMMA"]="UTCHMMA"; pat["LDTM"]="LDTM"; pat["UTMALDG"]="UTMALDG";
          pat["UTMASTG"]="UTMASTG"; pat["UTCBAR"]="UTCBAR"; pat["SYNCS"]="SYNCS"; pat["UTMAPF"]="UTMAPF|UTMACCTL" }
  /Function :/ { }
  END { for (k in c) { split(k, a, SUBSEP); print a[1], a[2], c[k] } }' | sort | awk '
  { if ($1 != last) { if (last != "") print line; line = $1 ":"; last = $1 } line = line " " $2 "=" $3 }
  END { print line }' | grep -E "UTCHMMA|UTMALDG|LDTM" | while read -r l; do
    name=$(echo "$l" | cut -d: -f1 | c++filt | sed 's/(.*//'); echo "$name:$(echo "$l" | cut -d: -f2-)"; done
