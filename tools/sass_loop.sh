#!/bin/bash
# sass_loop.sh <object-or-so> <mangled kernel name> <out.sass>: instruction lines of one kernel, one per line
cuobjdump -sass -fun "$2" "$1" | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's#/\* 0x[0-9a-f]+ \*/##; s/\s+$//' > "$3"
wc -l "$3"
