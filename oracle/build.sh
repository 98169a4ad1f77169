#!/bin/sh
# Builds the C oracle (test infrastructure) into oracle/_build/.  Called by __graft_entry__.build().
set -e
here="$(cd "$(dirname "$0")" && pwd)"
mkdir -p "$here/_build"
gcc -O2 -fPIC -shared -fopenmp -ffp-contract=off -Wall -Wextra -o "$here/_build/libnsm_oracle.so" "$here/nsm_oracle.c" -lm
echo "built $here/_build/libnsm_oracle.so"
