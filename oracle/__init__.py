"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference algorithm for the b200sr hot path.

Nothing in the product package imports this directory. It is used by tests/, by __graft_entry__.smoke() and by
bench.py's cpu_baseline / `--impl reference` leg, as the checker and as the timed CPU baseline.

Parity status: the UNet restatement (unet_oracle.py) is PINNED against the reference's own modules, imported
unmodified from /root/reference/src in the build container by oracle/make_golden.py, which also wrote the
committed fixtures in tests/golden/. The reference ships no tests or golden vectors of its own.
The SSIM part of the combined loss is "parity unpinned" against the reference (its source notebook is missing
from the snapshot); its uniform-window mode is pinned against a scipy restatement of skimage's
structural_similarity, the only SSIM the reference calls (src/VolumeVisualization.py:256).
"""
