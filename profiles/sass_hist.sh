#!/bin/bash
# sass_hist.sh : static SASS opcode histograms (cuobjdump -sass) of the hot kernels of the shipped library
lib=gmx-acqueduct_b200/libwn_b200.so
out=${1:-profiles/r02}
for k in "wn_k_edgesILb1ELi0ELi2ELb0E" "wn_k_copyILi0E" "wn_k_copyILi1E" "wn_k_cp_sweepILi0ELi0E" "wn_k_cp_sweepILi1ELi0E" "wn_k_scatter" "wn_k_bin"; do
  f=$out/sass_hist_$(echo $k | tr -c 'A-Za-z0-9_' '_').txt
  cuobjdump -sass -fun $(cuobjdump -elf $lib 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*$k[A-Za-z0-9_]*" | sort -u | head -1) $lib 2>/dev/null > /tmp/sass_$$.txt
  {
    echo "# static SASS of $(grep -m1 'Function :' /tmp/sass_$$.txt)"
    echo "# instructions: $(grep -cE '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_$$.txt)"
    grep -E '^\s+/\*[0-9a-f]{4}\*/' /tmp/sass_$$.txt | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/^@!?U?P[0-9T]+\s+//' | awk '{split($1,a,"."); print a[1]}' | sort | uniq -c | sort -rn
    echo "# Blackwell/Hopper-specific or notable opcodes:"
    grep -oE '\b(FFMA2|FMUL2|FADD2|UTMALDG|UBLKCP|LDGSTS|SYNCS|TCGEN05|UTCMMA|ATOMG|REDG|RED|MUFU\.RSQ64H|MUFU\.RCP64H|DFMA|DMUL|DADD|LDG\.E\.128[A-Z.]*|STG\.E\.128|LDS\.128)\b' /tmp/sass_$$.txt | sort | uniq -c | sort -rn
  } > $f
  echo "$f: $(sed -n 2p $f)"
done
rm -f /tmp/sass_$$.txt
