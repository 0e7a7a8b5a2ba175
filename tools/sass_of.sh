#!/bin/bash
# usage: sass_of.sh <lib.so> <mangled-substring>   -> plain SASS of that kernel (no line info) on stdout
set -e
T=$(mktemp -d); cd $T
cuobjdump -xelf all "$1" > /dev/null
C=$(ls -S *.cubin | head -1)
nvdisasm -c "$C" | awk -v k="$2" '/^\.text\./{on=index($0,k)>0} on{print}' 
rm -rf $T
