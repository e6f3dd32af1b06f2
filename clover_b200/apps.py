"""The reference's quantized linear-algebra applications, composed from the hot-path containers.

Host-side mirror of test/performance/01_measure.h: Q_IHT (:924-946, iterative hard thresholding for compressive
sensing) and Q_GD (:1000-1020, gradient descent). Every step is one C-ABI call on the current CUDA stream; nothing
is read back to the host inside the loop, so the whole iteration sequence can also be captured in a CUDA graph.
"""
from __future__ import annotations

from ._lib import THRESHOLD_AUTO


def Q_IHT(Phi, PhiT, x, y, t1, t2, t3, iterations: int, K: int, mu: float, threshold_mode: int = THRESHOLD_AUTO) -> None:
    """x <- H_K(x + mu * Phi^T (y - Phi x)), `iterations` times, all operands quantized (01_measure.h:924-946).

    Phi [M x N], PhiT [N x M] (= Phi.transpose), x / t3 length N, y / t1 / t2 length M. The reference calls the
    `_parallel` twins; here every method has one (GPU) implementation."""
    x.clear()
    for _ in range(int(iterations)):
        Phi.mvm(x, t1)                      # t1 = Phi * x
        y.scaleAndAdd(t1, -1.0, t2)         # t2 = y - Phi * x
        PhiT.mvm(t2, t3)                    # t3 = Phi' * (y - Phi * x)
        x.scaleAndAdd(t3, mu)               # x = x + mu * Phi' * (y - Phi * x)
        x.threshold(K, threshold_mode)      # hard thresholding


def Q_GD(Phi, PhiT, x, y, t1, t2, t3, iterations: int, mu: float) -> None:
    """Gradient descent on ||y - Phi x||^2 with quantized operands (01_measure.h:1000-1020)."""
    x.clear()
    for _ in range(int(iterations)):
        Phi.mvm(x, t1)
        y.scaleAndAdd(t1, -1.0, t2)
        PhiT.mvm(t2, t3)
        x.scaleAndAdd(t3, mu)


def capture(fn, warmup: int = 2):
    """Capture ``fn`` (a sequence of container calls with no host read-back, e.g. a Q_IHT call) into a CUDA graph.

    Kernel scratch (tickets, fp32 intermediates) is keyed by (device, stream) and allocated on first use, which is not
    allowed inside a capture: ``fn`` is therefore first run ``warmup`` times on the SAME side stream the capture then
    uses, so that every allocation has happened and the graph references that stream's blocks (which are never freed).
    Replaying the graph removes the per-call host cost - five launches per IHT iteration otherwise each pay a
    Python/ctypes round trip.

    Containers with a PRNG key cannot be captured: the kernels receive the key lanes BY VALUE, so every replay would
    reuse the same rounding noise and the host key would no longer track the stream the reference consumes
    (``_Keyed._key_ptr`` raises while a capture is in progress)."""
    import torch
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(warmup):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        fn()
    return graph
