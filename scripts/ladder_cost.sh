#!/bin/bash
# Static cost of the fast decoder's PTX blocks for a lone warp (scripts/sass_cost.py), from a single-kernel build:
# seconds per try, no GPU.  usage: scripts/ladder_cost.sh [extra nvcc flags]
set -e
D=$(mktemp -d)
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cat > $D/one.cu <<'EOC'
#include "lzgpu_prep.h"
#include "lzgpu_unit.cuh"
using namespace lzgpu;
struct KArgs { const lzgpu_unit *units; const int32_t *order; const uint8_t *in_base; uint8_t *out_base; lzgpu_result *results; uint16_t *lit_ws; uint64_t lit_ws_stride; uint32_t lit_bits_cap, slot0, stage_off; uint32_t *progress; };
template <int kV>
__global__ void __launch_bounds__(32, 14) k1(const KArgs a) {
    extern __shared__ __align__(16) uint16_t smem_probs[];
    const uint32_t slot = a.slot0 + blockIdx.x;
    const int32_t ui = a.order[slot];
    const lzgpu_unit u = a.units[ui];
    uint16_t *P = smem_probs, *L = smem_probs + LZ_LAY(kV)::LIT;
    UnitIO io; io.in = a.in_base + u.in_off; io.in_len = u.in_len; io.out = a.out_base + u.out_off; io.out_cap = u.out_cap;
    io.stage = reinterpret_cast<uint8_t *>(smem_probs + a.stage_off); io.inbuf = io.stage + 128; io.progress = a.progress ? a.progress + ui : nullptr; io.hout = nullptr; io.push_stat = nullptr;
    run_unit_lzma1<kV>(u, io, P, L, a.results[ui]);
}
template __global__ void k1<97>(const KArgs);
EOC
nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -I$ROOT/lzma_b200/csrc -cubin -o $D/one.cubin $D/one.cu -Xptxas -v "$@" 2>&1 | grep -i "registers\|error\|spill" || true
python3 $ROOT/scripts/sass_cost.py $D/one.cubin k1 --by-line lzgpu_fast2.cuh > $D/cost.txt
for m in "F2_BIT(d, p_im" "F2_LIT_PRE(d, sym" "F2_LIT(d, sym" "F2_ISREP_LEN(d, len" "F2_TREE6(d, slot" "F2_TREE4(d, m" "F2_LEN(d, len"; do
  ln=$(grep -n -F "$m" $ROOT/lzma_b200/csrc/lzgpu_fast2.cuh | tail -1 | cut -d: -f1)
  printf "%-22s line %4s: %s\n" "$m" "$ln" "$(grep ":$ln " $D/cost.txt | sed 's/.*cuh:[0-9]* *//')"
done
tail -2 $D/cost.txt | head -1
cp $D/one.cubin /tmp/last_one.cubin
rm -rf $D
