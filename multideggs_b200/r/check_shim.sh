#!/usr/bin/env bash
# Syntax/type check of the R `.Call` shim against stub R headers (R is not installed in the build image).
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
gcc -std=c11 -Wall -Wextra -Wno-unused-parameter -Wno-cast-function-type -fsyntax-only -I "$here/src/stub" -I "$here/../../include" "$here/src/rcall.c"
