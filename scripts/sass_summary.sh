#!/bin/bash
# Per-kernel counts of the SASS mnemonics that prove the Blackwell-native paths (tcgen05.mma -> UTCHMMA, tcgen05.ld -> LDTM,
# TMA -> UTMALDG, ex2 -> MUFU.EX2; HMMA = legacy mma.sync) in the shipped library, summed over template instances.
# usage: scripts/sass_summary.sh > profiles/r2_sass_summary.txt
SO=${1:-kvcache_factory_b200/libkvc_b200.so}
echo "cuobjdump -sass $SO, sm_100a; static instruction counts per kernel, all template instances summed"
echo "instances UTCHMMA   LDTM UTMALDG MUFU.EX2   HMMA LDGSTS  kernel"
cuobjdump -sass "$SO" | awk '/Function :/ { print "FN " $3; next } /UTCHMMA/ {print "A"} /LDTM/ {print "B"} /UTMALDG/ {print "C"} /MUFU.EX2/ {print "D"} /[^C]HMMA/ {print "E"} /LDGSTS/ {print "F"}' \
 | while read tag name; do if [ "$tag" = FN ]; then echo "FN $(echo "$name" | c++filt | sed -e 's/^void //' -e 's/<.*//' -e 's/(.*//')"; else echo "$tag"; fi; done \
 | awk '/^FN/ { fn=$2; inst[fn]++; next } { c[fn, $1]++ } END { for (k in inst) printf "%9d %7d %6d %7d %8d %6d %6d  %s\n", inst[k], c[k,"A"], c[k,"B"], c[k,"C"], c[k,"D"], c[k,"E"], c[k,"F"], k }' | sort -k8
