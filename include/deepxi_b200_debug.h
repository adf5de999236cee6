/*
 * deepxi_b200_debug.h -- entry points of the TUNING build only (libdeepxi_b200_dbg.so: DXI_DEBUG_BUILD=1 python -m
 * deepxi_b200.build, compiled with -DDXI_ENABLE_DEBUG).  None of these symbols, nor the kernel arguments they feed, exist in the
 * product library libdeepxi_b200.so.  Select the tuning build at run time with DXI_LIB=<path> (deepxi_b200/_lib.py); used by
 * scripts/chain_clocks.py, tcn_clocks.py, tcn_timeline.py, tmem_bw.py, tmem_layout.py.
 */
#ifndef DEEPXI_B200_DEBUG_H_
#define DEEPXI_B200_DEBUG_H_
#include "deepxi_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Debug aid for kernel tuning: the epilogue of stage `stage` of subsequent dxi_net_forward calls from this
 * thread writes 16 clock64 stamps per tile into dev_buf (int64 [n_tiles * 16]); NULL switches it off. */
DXI_API void dxi_debug_tcn_clocks(long long* dev_buf, int stage);

/* Debug aid: subsequent dxi_net_forward calls of this thread run only the stem and stages 0..stage (-1: all). */
DXI_API void dxi_debug_tcn_stop_after(int stage);

/* Tuning aid: cycles for `warps` warps x `rounds` x 4 KB tcgen05.ld (mode 0) / st (mode 1) / both (mode 2) on one SM;
 * dev_out[0] = cycles. */
DXI_API int dxi_debug_tmem_bw(int mode, int warps, int rounds, long long* dev_out, void* stream);
/* Register <-> (lane, column) map of tcgen05.ld.16x256b.x8 (dev_out[0 .. 128*64)) and tcgen05.st.16x128b.x8 (dev_out[128*64 .. +128*32)),
 * each entry lane * 256 + column; scripts/tmem_layout.py checks them against the closed forms quoted in umma.cuh. */
DXI_API int dxi_debug_tmem_layout(uint32_t* dev_out, void* stream);

/* Depth-first ResNetV2 kernel (tcn_chain.cu): the epilogue's thread 0 and the MMA warp of the CTA that processes work item `item` of
 * subsequent forwards write clock64 stamps into dev_buf (int64 [n_blocks * 32]); NULL switches it off. */
DXI_API int dxi_debug_chain_clocks(long long* dev_buf, int item);

#ifdef __cplusplus
}
#endif
#endif /* DEEPXI_B200_DEBUG_H_ */
