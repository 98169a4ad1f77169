#!/bin/sh
# Blackwell evidence from the built library: per cubin, the SASS mnemonics that show what the
# kernels are made of (TMA bulk copies + mbarrier, byte-wise SAD, popcount, carry-chain adds,
# votes / shuffles / reductions, 64-bit shared-memory atomics).  Usage: tools/sass_summary.sh [lib.so]
LIB=${1:-napkon-string-matching_b200/napkon_string_matching/gpu/libnsm_b200.so}
TMP=$(mktemp -d)
( cd "$TMP" && cuobjdump -xelf all "$OLDPWD/$LIB" > /dev/null )
echo "# cuobjdump -sass of $LIB ($(nvcc --version | tail -1))"
echo "# counts of selected mnemonics per cubin (arch sm_100a); no tcgen05 / UTCMMA / HMMA by design:"
echo "# the path is integer set / bit-vector work with no dense contraction (BASELINE.json north_star)"
for f in jaccard qratio qratio_flat qratio_long pack microbench; do
  c="$TMP/$f.sm_100a.cubin"
  [ -f "$c" ] || continue
  echo "== $f.sm_100a.cubin  (kernels: $(cuobjdump -sass "$c" | grep -c 'Function :'))"
  cuobjdump -sass "$c" | grep -oE '^\s+/\*[0-9a-f]+\*/\s+(@!?U?P[0-9T] )?[A-Z0-9_.]+' | awk '{print $NF}' |
    grep -E '^(UBLKCP|SYNCS|VABSDIFF4|POPC|LOP3|IADD3|VOTE|SHFL|REDUX|ATOMS|ATOMG|RED|LDS|STS|LDG|STG|DFMA|DMUL|DADD|MUFU|BAR|UTCMMA|HMMA|IMMA|UTMALDG)' |
    sed -E 's/^(LDS|STS|LDG|STG|LOP3|IADD3|POPC|SHFL|VOTE|REDUX|ATOMS|ATOMG|RED|BAR|DFMA|DMUL|DADD|MUFU)(\..*)?$/\1/' |
    sort | uniq -c | sort -rn | awk '{printf "  %8d %s\n", $1, $2}'
done
rm -rf "$TMP"
