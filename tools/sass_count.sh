#!/bin/bash
# static SASS instruction count per kernel of a built library: tools/sass_count.sh lib.so [pattern]
cuobjdump -sass "$1" | awk '/Function :/{name=$3} /^[ \t]+\/\*[0-9a-f][0-9a-f][0-9a-f][0-9a-f]\*\//{cnt[name]++} END{for(n in cnt) print cnt[n], n}' | sort -n | grep -E "${2:-.}"
