"""CPU restatement (oracle) of the Deep Xi inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``deepxi_b200/`` may import this package;
only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs use it, and there only as the checker / the timed CPU arm.

The reference (golfbears/DeepXi) cannot be imported here (TensorFlow, tensorflow_addons,
librosa, soundfile are absent and ``deepxi/model.py:30-31`` imports packages that are not
in its tree), so every function below restates the reference's arithmetic in numpy /
scipy / torch-CPU and cites the reference file:line it follows.

Parity pinning (see tests/test_oracle_kat.py and DESIGN.md):
  * STFT -> MMSE-LSA -> iSTFT -> int16 is pinned to the reference's shipped output
    (out/resnet-1.0c/e180/...), max |delta| = 1 LSB.
  * ResNetV2 / MHANetV3 forward, the inverse CDF map alone and the non-LSA gains have no
    reference artefact that pins them numerically ("parity unpinned" for those; they are
    cross-checked against independent implementations: scipy.special, torch.nn.functional).
"""
from . import sig, cdfmap, gain, tcn, attention, wavio, pipeline, train_tgt  # noqa: F401
