"""B200-native QuadtreeCNN hot path: CUDA kernels + C ABI (csrc/, libqtcnn.so) and the host-side mirror of the
reference's `models.py` interface. Import through the `qtcnn_b200` alias package."""
