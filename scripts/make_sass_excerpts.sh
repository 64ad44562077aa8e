#!/bin/bash
# profiles/r2_sass_excerpts.md: what the built library's SASS says about the two hot kernels (run after `make lib`)
LIB=paris_b200/libparis_b200.so
OUT=profiles/r2_sass_excerpts.md
TMP=$(mktemp)
cuobjdump -sass $LIB > $TMP
{
echo "# SASS evidence, round 2 (\`cuobjdump -sass $LIB\`, nvcc 12.9, sm_100a)"
echo
echo "Counts over the whole library:"
echo
echo "| mnemonic | count | what it is |"
echo "|---|---|---|"
for m in "UTMALDG.3D:TMA tensor load, plain stack layout (cp.async.bulk.tensor.3d)" "UTMALDG.4D:TMA tensor load, parity-split layout (cp.async.bulk.tensor.4d)" "SYNCS:mbarrier operations (init / arrive.expect_tx / try_wait)" "FFMA2:packed f32x2 fused multiply-add" "FMUL2:packed f32x2 multiply" "FADD2:packed f32x2 add" "LDS:shared-memory loads" "HMMA:legacy tensor-core MMA (must be 0)" "UTCHMMA:tcgen05 MMA (0: gather + interpolation, no contraction)"; do
  k=${m%%:*}; d=${m#*:}
  echo "| \`$k\` | $(grep -c "$k" $TMP) | $d |"
done
echo
for fn in "bp_tma_kernelINS_8tile_cfgILi8ELi8ELi4ELi8ELi32ELi312ELi2ELb1EEELb0" "filter_kernelILi12ELb1"; do
  start=$(grep -n "Function : .*$fn" $TMP | head -1 | cut -d: -f1)
  [ -z "$start" ] && continue
  end=$(awk -v s=$start 'NR>s && /Function :/ {print NR; exit}' $TMP)
  [ -z "$end" ] && end=$(wc -l < $TMP)
  sed -n "${start},${end}p" $TMP | grep -v "^\s*/\* 0x" > $TMP.fn
  echo "## \`$(sed -n 1p $TMP.fn | sed 's/.*Function : //')\`"
  echo
  echo "Instruction mix (top 16):"
  echo
  echo '```'
  awk '/^ +\/\*[0-9a-f]+\*\//{print $2}' $TMP.fn | sed 's/;//' | sed 's/\..*//' | sort | uniq -c | sort -rn | head -16
  echo '```'
  echo
  echo "Excerpts (TMA issue, mbarrier wait, packed FP32, gathers):"
  echo
  echo '```'
  grep -E "UTMALDG|SYNCS" $TMP.fn | head -6
  grep -E "FFMA2|FMUL2|FADD2" $TMP.fn | head -6
  grep -E " LDS " $TMP.fn | head -6
  echo '```'
  echo
done
} > $OUT
rm -f $TMP $TMP.fn
echo written $OUT
