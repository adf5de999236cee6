"""deepxi_b200: B200-native implementation of the Deep Xi inference hot path.

noisy waveform -> STFT -> a priori SNR estimator (ResNetV2 TCN / MHANetV3) -> inverse CDF map
-> gain function -> iSTFT, behind the Python API surface of golfbears/DeepXi
(deepxi.sig, deepxi.map, deepxi.gain, deepxi.inp_tgt, deepxi.network.selector, deepxi.model).
All arithmetic runs in hand-written sm_100a CUDA kernels reached through the C ABI of
libdeepxi_b200.so (include/deepxi_b200.h); there is no CPU fallback.
"""
__version__ = '0.1.0'
