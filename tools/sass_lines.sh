#!/bin/bash
# static SASS instruction count per source line of one kernel: sass_lines.sh obj.o kernel [top]
set -e
d=$(mktemp -d); cp "$1" $d/o.o; (cd $d && cuobjdump -xelf all o.o >/dev/null)
nvdisasm --print-line-info $d/*.cubin 2>/dev/null | awk -v K="$2" '
/^\/\/-+ \.text\./ {infn = ($0 ~ ("\\.text\\." K " "))}
/\/\/## File/ {loc=$0; sub(/.*\//,"",loc); sub(/ inlined.*/,"",loc); gsub(/"/,"",loc)}
infn && /^ +\/\*[0-9a-f]+\*\// {cnt[loc]++; tot++}
END {for (l in cnt) print cnt[l], l; print tot, "TOTAL"}' | sort -rn | head -${3:-40}
rm -rf $d
