#!/bin/bash
# static SASS statistics of one kernel in libvxgpu.so: instruction count and opcode histogram
# usage: tools/sass_stats.sh <mangled-name-substring>
LIB=${LIB:-evosoro_b200/lib/libvxgpu.so}
cuobjdump -sass -fun "$1" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//' | sed -E 's/^@!?U?P[0-9T]+\s+//' | awk '{print $1}' | sed 's/\..*//;s/;//' | sort | uniq -c | sort -rn | awk '{t+=$1; print} END {print t, "TOTAL"}' | (head -25; echo ...; tail -1)
