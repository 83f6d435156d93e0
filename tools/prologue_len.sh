#!/bin/bash
# Static check: SASS instruction count of a kernel and of its part before the first backward branch target
# (prologue) -- usage: tools/prologue_len.sh <lib.so> <mangled-kernel-substring>
cuobjdump -sass "$1" | awk -v pat="$2" '
  /Function :/ { on = index($0, pat) > 0; if (on) print $0 }
  on && /\/\*[0-9a-f]+\*\// { n++ }
  END { print "instructions:", n }'
