#!/bin/bash
# usage: tools/sass_of.sh <kernel-name-substring> : plain SASS listing (one instruction per line) of one kernel of libema_b200.so
cuobjdump -sass ema_b200/libema_b200.so | awk -v k="$1" '/Function : /{f=($0 ~ k)} f' | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#^\s*/\*[0-9a-f]*\*/\s*##; s#\s*/\*[0-9a-fx]*\*/##g'
