#!/bin/bash
# md5 of the SASS (addresses and encodings stripped) of every object of the library: two builds with the
# same fingerprint run the same device code.  Used to check that compiled-out experiments and refactors
# leave the validated kernels untouched.   usage: tools/sass_fingerprint.sh [csrc dir]
d=${1:-diskann_b200/csrc}
for o in "$d"/*.o; do
  h=$(cuobjdump -sass "$o" | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed 's#/\* 0x[0-9a-f]* \*/##' | md5sum | cut -c1-16)
  echo "$(basename "$o" .o) $h"
done
