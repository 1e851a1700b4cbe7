#!/bin/bash
# Build one library per forward-schedule variant into .ab/ (same-box A/B with scripts/ab.sh).
#   scripts/ab_build.sh name "-DOMNI_V_DDIAG=1 -DOMNI_V_XMODE=3" [name2 "flags2" ...]
set -e
cd "$(dirname "$0")/.."
mkdir -p .ab
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  touch omnimamba_b200/csrc/ssd_tc.cu
  OMNI_NVCC_EXTRA="$flags" python -m omnimamba_b200.build > /dev/null
  cp omnimamba_b200/lib/libomnissm.so .ab/lib_$name.so
  grep -A2 "Compiling entry function.*ssd_tc_fwd_kernelILi0ELb0" omnimamba_b200/build/ptxas.log | grep -o "Used [0-9]* registers\|[0-9]* bytes spill stores" | tr '\n' ' '
  echo " <- $name ($flags)"
done
touch omnimamba_b200/csrc/ssd_tc.cu
python -m omnimamba_b200.build > /dev/null   # back to the default build
