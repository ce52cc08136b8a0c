"""Host <-> device pipelining for the projection step (public helper).

A training step of the reference feeds the projection from host data (the
DataLoader's batch) and reads results back (losses, gradients for logging or a
host-side optimiser).  Done naively on one stream, the copies serialise with
the kernels: at workload A (64 projections) the step moves 6 MB in and 8 MB out
over PCIe for ~0.2 ms of kernels.  ``HostPipeline`` is the usual three-stream
prefetcher: step i's inputs are copied in on a copy stream while step i-1
computes, and step i-1's results are copied out on a third stream while step i
computes.  Every step still copies ITS inputs from pinned host memory and ITS
results to pinned host memory; only the waiting is overlapped.

    pipe = HostPipeline(device, depth=3)
    for batch in loader:                       # dict of pinned CPU tensors
        dev = pipe.upload(batch)               # async H2D on the copy stream
        out = step_fn(dev)                     # user code on the current stream
        pipe.download(out, host_buffers)       # async D2H on the readback stream
    pipe.drain()

The kernels are untouched: this only orders copies and computation with CUDA
events, and recycles ``depth`` sets of device input buffers so that no
allocation or synchronisation happens in steady state.
"""
import os

import torch


def bind_to_device_numa(device_index):
    """Pin the calling process to the CPU cores local to GPU ``device_index`` (NVML's CPU
    affinity of the device), so that pinned host buffers allocated afterwards -- and the copies
    that read and write them -- stay on the GPU's own NUMA node.  With one process per GPU on a
    two-socket host, an unbound rank whose staging buffers land on the far socket shares the
    inter-socket link with every other rank's copies.  Returns the CPU set, or None when NVML or
    the affinity call is unavailable (the process is left as it was)."""
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            tag = ids[device_index]
            handle = (pynvml.nvmlDeviceGetHandleByUUID(tag) if tag.startswith(("GPU-", "MIG-"))
                      else pynvml.nvmlDeviceGetHandleByIndex(int(tag)))
        else:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        n_cpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (n_cpu + 63) // 64)
        cpus = {64 * w + b for w, word in enumerate(words) for b in range(64) if (word >> b) & 1}
        cpus &= set(os.sched_getaffinity(0))
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return cpus
    except Exception:
        return None


class HostPipeline:
    def __init__(self, device, depth=3):
        if depth < 2:
            raise ValueError("depth must be >= 2 (one set in flight, one being filled)")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("HostPipeline needs a CUDA device: dpc_b200 has no CPU fallback")
        self.depth = depth
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._slots = [dict() for _ in range(depth)]        # name -> device tensor
        self._free = [None] * depth                          # event: slot's last consumer done
        self._out_done = [None] * depth                      # event: slot's D2H finished
        self._ready = [torch.cuda.Event() for _ in range(depth)]   # events are reused, not rebuilt
        self._done = [torch.cuda.Event() for _ in range(depth)]
        self._out = [torch.cuda.Event() for _ in range(depth)]
        self._i = 0
        self.h2d_bytes = 0
        self.d2h_bytes = 0

    def upload(self, host_tensors):
        """Copy a dict of pinned CPU tensors into this step's device buffers on the
        copy stream; the current stream waits for the copy (not the host)."""
        cur = torch.cuda.current_stream(self.device)
        k = self._i % self.depth
        slot = self._slots[k]
        if self._free[k] is not None:
            self.h2d.wait_event(self._free[k])               # buffers still read by step i-depth
        n = 0
        with torch.cuda.stream(self.h2d):
            for name, t in host_tensors.items():
                if t is None:
                    slot[name] = None
                    continue
                if not t.is_pinned():
                    raise ValueError("%s: host tensors must be pinned for asynchronous copies" % name)
                d = slot.get(name)
                if d is None or d.shape != t.shape or d.dtype != t.dtype:
                    if d is not None:
                        # the replaced buffer was allocated on the copy stream but read by
                        # kernels on the compute stream: tell the allocator before dropping it
                        d.record_stream(cur)
                    d = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                    slot[name] = d
                d.copy_(t, non_blocking=True)
                n += t.numel() * t.element_size()
            ready = self._ready[k]
            ready.record(self.h2d)
        cur.wait_event(ready)
        self.h2d_bytes = n
        self._cur = k
        return dict(slot)

    def download(self, device_tensors, host_buffers):
        """Copy results to pinned host buffers on the readback stream, after the
        current stream has produced them; marks this step's input buffers free."""
        cur = torch.cuda.current_stream(self.device)
        k = self._cur
        done = self._done[k]
        done.record(cur)
        self._free[k] = done
        self.d2h.wait_event(done)
        n = 0
        with torch.cuda.stream(self.d2h):
            for name, t in device_tensors.items():
                h = host_buffers[name]
                h.copy_(t.detach(), non_blocking=True)
                t.record_stream(self.d2h)
                n += t.numel() * t.element_size()
            out = self._out[k]
            out.record(self.d2h)
        self._out_done[k] = out
        self.d2h_bytes = n
        self._i += 1
        return out                                           # wait on it before reading host_buffers

    def fork(self):
        """Start of a CUDA-graph capture on the current stream: forget events recorded
        outside the capture and bring both copy streams into it."""
        cur = torch.cuda.current_stream(self.device)
        self._free = [None] * self.depth
        self._i = 0
        self.h2d.wait_stream(cur)
        self.d2h.wait_stream(cur)

    def join(self):
        """End of a capture (or of a timed region): the current stream waits for both
        copy streams."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.h2d)
        cur.wait_stream(self.d2h)

    def drain(self):
        """Block until every outstanding copy has landed."""
        self.h2d.synchronize()
        self.d2h.synchronize()
        torch.cuda.current_stream(self.device).synchronize()


class GraphedSteps:
    """``n_steps`` consecutive calls of ``step_fn(k)`` captured ONCE into a CUDA graph.

    At ~0.2 ms of kernels per step the eager path is bound by the host (Python,
    the autograd engine's thread hand-off, ~30 launches and copies per step);
    replaying a captured graph costs one launch for ``n_steps`` steps.  The
    step function is ordinary user code -- ``HostPipeline.upload`` from fixed
    pinned staging buffers, ``pointcloud_project_fast``, a loss,
    ``torch.autograd.grad`` / ``backward``, ``HostPipeline.download`` -- and the
    copies of step k overlap the kernels of steps k-1 / k+1 inside the graph.
    Shapes, pointers of the staging buffers and the set of ops must not change
    between replays (write each new batch INTO the staging buffers).
    """

    def __init__(self, step_fn, n_steps, device, pipe=None, warmup=2):
        self.device = torch.device(device)
        self.n_steps = int(n_steps)
        cur = torch.cuda.current_stream(self.device)
        for _ in range(warmup):                      # allocator, lazy init, cached scratch
            for k in range(self.n_steps):
                step_fn(k)
        if pipe is not None:
            pipe.drain()
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):
            with torch.cuda.graph(self.graph, stream=side):
                if pipe is not None:
                    pipe.fork()
                for k in range(self.n_steps):
                    step_fn(k)
                if pipe is not None:
                    pipe.join()
        cur.wait_stream(side)
        if pipe is not None:
            pipe.fork()                              # leave no captured events behind

    def replay(self):
        """Enqueue all ``n_steps`` steps on the current stream."""
        self.graph.replay()


class AlternatingGraphs:
    """``lanes`` independent ``GraphedSteps`` (each with its own ``HostPipeline`` and buffers),
    replayed in turn on ``lanes`` private streams.

    One captured graph is a closed unit: its first host-to-device copy has nothing to overlap and
    neither has its last device-to-host copy, and two replays on one stream run back to back, so
    every replay drains the copy pipeline (measured at workload A: ~340 us per replay).  With two
    lanes the next replay's first steps run under the previous replay's tail.

    ``make_step(lane, pipe)`` returns the ``step_fn(k)`` of that lane; it must use the given
    pipeline and lane-private host / device buffers."""

    def __init__(self, make_step, n_steps, device, lanes=2, warmup=2):
        self.device = torch.device(device)
        self.streams = [torch.cuda.Stream(self.device) for _ in range(lanes)]
        self.pipes = [HostPipeline(self.device) for _ in range(lanes)]
        self.lanes = [GraphedSteps(make_step(l, self.pipes[l]), n_steps, self.device,
                                   pipe=self.pipes[l], warmup=warmup) for l in range(lanes)]
        self.n_steps = int(n_steps)
        self._i = 0

    def replay(self):
        """Enqueue the next lane's ``n_steps`` steps, ordered after the work already on the current
        stream (and after that lane's previous replay)."""
        l = self._i % len(self.lanes)
        self._i += 1
        s = self.streams[l]
        s.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(s):
            self.lanes[l].replay()

    def join(self):
        """The current stream waits for every lane."""
        cur = torch.cuda.current_stream(self.device)
        for s in self.streams:
            cur.wait_stream(s)
