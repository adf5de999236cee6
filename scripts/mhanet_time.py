"""Times the MHANetV3 forward (config C3: 64 utterances x 30 s, T = 1875) with CUDA events (tuning aid)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepxi_b200 import weights, _lib
from deepxi_b200.network.selector import network_selector
B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 1875
kw = dict(d_model=256, n_blocks=5, n_heads=8, warmup_steps=40000, max_len=2048, causal=1, outp_act='Sigmoid')
prec = sys.argv[3] if len(sys.argv) > 3 else 'f32'
net = network_selector('MHANetV3', None, 257, precision=prec, **kw).load_weights(weights.synthetic_mhanetv3(0))
x = torch.rand(B, T, 257, device='cuda')
for _ in range(2):
    y = net(x)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
_lib.launch_count_reset()
e0.record()
n = 3
for _ in range(n):
    y = net(x)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
flop = B * T * 17.73e6
print(prec, 'MHANetV3 %d x %d frames: %.2f ms / forward, %.1f TFLOP/s (17.73 MFLOP/frame unmasked), %.0f audio-s/s, %d launches / forward'
      % (B, T, ms, flop / ms / 1e9, B * T * 0.016 / (ms / 1e3), _lib.launch_count() // n))
_lib.profile_enable(True)
for k in ('mha_gemm', 'mha_attn'):
    _lib.profile_read(k)
y = net(x); torch.cuda.synchronize()
for k in ('mha_gemm', 'mha_attn'):
    ms_k, n_k = _lib.profile_read(k)
    print('  %-9s %.2f ms in %d launches' % (k, ms_k, n_k))
_lib.profile_enable(False)
