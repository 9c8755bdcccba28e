#!/bin/bash
# A/B of attention-kernel library variants: tools/gpu/attn_ab.sh "<lib suffix or default>[:ENV=V,ENV=V]" ...
for spec in "$@"; do
  v=${spec%%:*}; envs=""; [ "$spec" != "$v" ] && envs=${spec#*:}
  lib=$PWD/mojo_opset_b200/libmojo_b200.so
  [ "$v" != "default" ] && lib=$PWD/mojo_opset_b200/libmojo_b200_$v.so
  echo "=== $spec"
  MOJO_B200_LIB=$lib python tools/attn_sweep.py $envs 2>&1 | tail -3
done
