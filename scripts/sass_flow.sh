#!/bin/bash
# usage: scripts/sass_flow.sh <object> <mangled kernel>  -- control skeleton of a kernel (loads, waits, barriers, MUFU runs)
cuobjdump -sass -fun "$2" "$1" | grep -v "^\s*$" > /tmp/flow.sass
grep -n "LDTM\|TRYWAIT\|BAR\|MUFU\|NOP" /tmp/flow.sass | awk -F: '{print $1, $2}' | awk '{print $1, $3, $4, $5}' | awk 'BEGIN{m=0} /MUFU/{m++; next} {if(m) print "   ...",m,"MUFU"; m=0; print}'
