#!/bin/bash
# usage: tools/sass_mix.sh <object-or-so> <kernel-name-regex>   -> static SASS opcode mix of the matching kernels
f=$1; pat=$2
cuobjdump -sass "$f" | awk -v pat="$pat" '
/Function :/ { on = ($0 ~ pat); if (on) { name=$3; print "== " name } ; next }
on && $1 ~ /^\/\*[0-9a-f]+\*\/$/ { op=$2; if (op ~ /^@/) op=$3; sub(/\..*/, "", op); sub(/;/, "", op); cnt[op]++; tot++ }
END { for (k in cnt) print cnt[k], k; print tot, "TOTAL" }' | sort -k1,1nr
