// TEST INFRASTRUCTURE ONLY.  Runs ba_pair_sums_kernel (onepiece_b200/csrc/opb_ba.cu, the part inside namespace opb, cut out by
// tests/test_ba_cpu.py into ba_device.inc) on the host emulator in cuda_emu.h.
#include "opb_common.cuh"
#include "ba_device.inc"

using namespace opb;

extern "C" void emu_ba_sums(const float *poses_rm, const int *src_id, const int *tgt_id, const long long *offset, int n_corr, const float *a,
                            const float *b, double *sums)
{
    emu::launch(n_corr, kBaThreads, [&] { ba_pair_sums_kernel(poses_rm, src_id, tgt_id, offset, a, b, sums); });
}
