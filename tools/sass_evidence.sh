#!/bin/bash
# SASS opcode histograms of the hot kernels (cuobjdump -sass over the objects of libgsb200.so) -> profiles/<tag>_sass_opcodes.md
# Evidence for: TMA bulk copies (UBLKCP) + mbarrier (SYNCS) in the Onesweep pass and the coarse-bin blend, packed fp32
# (FMUL2 / FADD2 / FFMA2) in the blend walk, no tensor-core opcodes anywhere (no dense contraction on this path).
cd "$(dirname "$0")/.."
tag=${1:-r2}
out=profiles/${tag}_sass_opcodes.md
hist() {  # object, mangled-name fragment, title
  echo "## $3"; echo
  cuobjdump -sass "$1" | awk -v K="$2" '/Function :/{on=index($0,K)>0} on && $1 ~ /^\/\*[0-9a-f]+\*\/$/ {op=$2; if (op ~ /^@/) op=$3; sub(/\..*/,"",op); sub(/;.*/,"",op); print op}' \
    | sort | uniq -c | sort -rn | awk '{printf "%s %s, ", $2, $1} END{print ""}'
  echo
}
{
  echo "# SASS opcode histograms ($tag; static counts over the whole kernel, cuobjdump -sass, sm_100a)"; echo
  hist 3dgs.cpp_b200/csrc/gsb_blend.o "k_blend2ILi0ELb0ELb1" "k_blend2<EXACT, no stats, COARSE> -- the kernel bench.py times (tile_cull 2)"
  hist 3dgs.cpp_b200/csrc/gsb_blend.o "k_blend2ILi0ELb0ELb0" "k_blend2<EXACT, no stats, per-tile lists>"
  hist 3dgs.cpp_b200/csrc/gsb_sort.o "k_onesweep_passIj" "k_onesweep_pass<u32> -- depth sort and instance sort"
  hist 3dgs.cpp_b200/csrc/gsb_preprocess.o "13k_emit_coarse" "k_emit_coarse"
  hist 3dgs.cpp_b200/csrc/gsb_preprocess.o "k_projectILb0ELb0" "k_project<no debug, not routed>"
  hist 3dgs.cpp_b200/csrc/gsb_preprocess.o "k_projectILb0ELb1" "k_project<ROUTED> -- frame sharding: survivors stored into peer memory"
  echo "## tensor-core / wgmma opcodes in libgsb200.so"; echo
  echo "HMMA/IMMA/UTCMMA/QGMMA lines: $(cuobjdump -sass 3dgs.cpp_b200/libgsb200.so | grep -c -E 'HMMA|IMMA|UTCHMMA|UTCQMMA|QGMMA|HGMMA')"
} > $out
echo wrote $out
