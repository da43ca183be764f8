#!/bin/bash
# SASS evidence for profiles/: how often the tcgen05 / TMA / tensor mnemonics occur in the built library, in total and per kernel.
#   scripts/sass_counts.sh > profiles/r2_sass_counts.txt
set -u
SO=${1:-eagleeverything_b200/libeaglegpu.so}
echo "# cuobjdump -sass $SO | grep -c <mnemonic>   (sm_100a; built by eagleeverything_b200/csrc/Makefile)"
TMP=$(mktemp)
cuobjdump -sass "$SO" > "$TMP"
for m in "UTCIMMA" "UTCIMMA.2CTA" "LDTM" "UTMALDG" "UBLKCP" "DMMA" "IDP.4A" "UTCBAR" "SYNCS" "MUFU.RCP64H" "DFMA"; do
  printf "%-14s %s\n" "$m" "$(grep -c -- "$m" "$TMP")"
done
echo
echo "# per kernel (demangled name up to the argument list): mnemonic count"
awk '/Function :/ {f=$3} /UTCIMMA|LDTM|UTMALDG|UBLKCP|DMMA|IDP\.4A/ {
        m=""; if ($0 ~ /UTCIMMA\.2CTA/) m="UTCIMMA.2CTA"; else if ($0 ~ /UTCIMMA/) m="UTCIMMA"; else if ($0 ~ /LDTM/) m="LDTM";
        else if ($0 ~ /UTMALDG/) m="UTMALDG"; else if ($0 ~ /UBLKCP/) m="UBLKCP"; else if ($0 ~ /DMMA/) m="DMMA"; else m="IDP.4A";
        c[f" "m]++ }
     END {for (k in c) print k, c[k]}' "$TMP" | while read f m c; do printf "%-70s %-14s %s\n" "$(echo $f | c++filt | sed 's/(.*//')" "$m" "$c"; done | sort
rm -f "$TMP"
