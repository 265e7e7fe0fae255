#!/bin/bash
# debug build of the library with the tcgen05 attention kernel's per-tile clock64 trace enabled for one CTA (-DT6_TRACE=<block>)
exec "$(dirname "$0")/build_variant_lib.sh" trace attention_tc.cu -DT6_TRACE=${1:-25001}
