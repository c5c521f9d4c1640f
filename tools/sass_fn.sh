#!/bin/sh
# usage: tools/sass_fn.sh <substring of mangled kernel name> > out.sass   (plain SASS of one kernel of the built library)
LIB=omega_match_b200/lib/libomega_match.so
FN=$(cuobjdump -sass $LIB 2>/dev/null | grep "Function :" | grep "$1" | head -1 | sed 's/.*Function : //')
cuobjdump -sass -fun "$FN" $LIB 2>/dev/null | grep -E "^\s+/\*[0-9a-f]{4,5}\*/" | sed -E 's/^\s+//; s/\s+\/\* 0x[0-9a-f]+ \*\/$//'
