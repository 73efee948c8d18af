"""CUDA-graph capture of a whole QAT step.

The steady-state step -- student forward, float-teacher forward, KD SI-SDR loss, backward, gradient gather,
all-reduce, global-norm clip, Adam -- is ~650 kernel launches with no host decision in between (observers off, no
`.item()`, the Adam step count lives on the device), so it can be captured once and replayed: the host cost of a
step drops from ~23 ms of Python/ctypes launch work to one `cudaGraphLaunch`, and kernel-to-kernel gaps shrink to
the graph's dependency latency.  Plumbing only; every node of the graph is a libfqss_sm100 kernel (plus NCCL's
all-reduce at world size > 1 and the memsets / copies the library enqueues).
"""
import torch

from . import roofline


class GraphedStep:
    """Capture `step_fn(*inputs) -> tensor(s)` after `warmup` eager runs; call the object with new inputs to replay.

    Inputs are copied into static buffers (same shapes / dtypes every call); outputs are the static tensors the
    captured step wrote (read them, or copy them out, before the next replay)."""

    def __init__(self, step_fn, example_inputs, warmup=3):
        if not all(t.is_cuda for t in example_inputs):
            raise RuntimeError("GraphedStep needs CUDA tensors (no CPU fallback)")
        self.static_in = [t.clone() for t in example_inputs]
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):               # warm up on a side stream: allocator pools, func attributes, caches
            for _ in range(max(warmup, 1)):
                step_fn(*self.static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        c0 = roofline.launch_count()
        with torch.cuda.graph(self.graph):
            self.static_out = step_fn(*self.static_in)
        self.kernels_per_replay = roofline.launch_count() - c0      # library kernels captured = launched by every replay

    def __call__(self, *inputs):
        for dst, src in zip(self.static_in, inputs):
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        self.graph.replay()
        from . import parallel
        parallel.bump_param_generation()      # a replayed step may contain an optimizer update (see parallel.param_generation)
        return self.static_out
