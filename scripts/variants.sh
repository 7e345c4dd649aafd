#!/bin/bash
# build tuning variants here (CPU), run them on the GPU box with scripts/run_variants.sh
# usage: bash scripts/variants.sh "TAG:-Dflags ..." ...
for v in "$@"; do tag=${v%%:*}; defs=${v#*:}; RLR_VARIANT=$tag RLR_DEFS="$defs" python -m rl_rocket_b200.csrc.build --force >/dev/null || echo "build of $tag failed"; done
ls -la rl_rocket_b200/_C/*.so
