"""CUDA-graph capture of a control step.

Every entry point of the C ABI (include/dust_b200.h) allocates nothing and only enqueues kernels on the stream it is
given, so a sequence of them -- e.g. one dual control step of the reference's drivers (particle_example.py:177-207,
simulations.py:104-138: SVMPC.optimize + forward, plant step, MPF.optimize) -- can be captured ONCE and replayed: the
launch-bound demo shapes lose their per-launch host cost (0.14 -> 0.10 ms per dual step, bench_configs.py).
Inputs and state must live in tensors that exist before the capture and are updated in place (the library's
`SvmpcCore`, `ops.mpf_optimize(x=...)` and the noise buffers are); values passed by value (seeds, learning rates) are
frozen into the graph."""
import torch


class CapturedStep:
    """`step = CapturedStep(fn)`; `step()` replays what `fn()` enqueued.  `fn` runs a few times eagerly first (on a side
    stream, as capture requires), so it must be a steady-state step that may be repeated."""

    def __init__(self, fn, warmup=3):
        if not torch.cuda.is_available():
            raise RuntimeError("dust_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                fn()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.result = fn()

    def __call__(self):
        self.graph.replay()
        return self.result
