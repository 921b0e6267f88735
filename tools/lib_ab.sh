#!/bin/bash
# tools/lib_ab.sh "<command>" lib_a.so lib_b.so ...: the same command under each build of the library (CLIPPPO_LIB), interleaved twice
cmd=$1; shift
for rep in 0 1; do for lib in "$@"; do echo "== [$rep] $lib"; CLIPPPO_LIB=$PWD/$lib bash -c "$cmd"; done; done
