#!/bin/bash
# tools/sass_of.sh <object> <substring of the mangled kernel name> -> SASS of that kernel (one instruction per line) on stdout
obj=$1; pat=$2
name=$(cuobjdump -elf "$obj" 2>/dev/null | grep -o "_Z[A-Za-z0-9_]*" | grep "$pat" | grep -v "_param_" | sort -u | head -1)
cuobjdump -sass -fun "$name" "$obj" | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed 's/\/\* 0x[0-9a-f]* \*\///'
