"""End-to-end (host-buffer) driver for CDR::run.

A caller whose tracer fields live in HOST memory runs one CDR step as: copy the
step's inputs to the device, set_rhom/set_Qm, run(), get_Qm, copy the result back.
Done naively that serialises ~40 B/update over PCIe around a ~ms kernel, so this
class tiles the tracers into chunks and pipelines chunk k's H2D copy, chunk k-1's
kernels and chunk k-2's D2H copy on separate CUDA streams (tracers are
independent problems, so chunking does not change any result). It uses only the
public CDR API of compose_b200 (declare/finish_setup once, then bulk
set_Qm/run/get_Qm per chunk).
"""
import os

import torch

import compose_b200 as cb


def cpus_near_gpu(device_index):
    """The CPUs NVML reports as local to a CUDA device (its NUMA node), intersected with
    what this process may run on; None when that cannot be determined."""
    try:
        import pynvml
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(device_index)
        try:
            bus = "%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
            h = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        nwords = (os.cpu_count() + 63)//64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, nwords)
        cpus = {64*i + b for i, w in enumerate(mask) for b in range(64) if (int(w) >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


class near_gpu:
    """Context manager: run the calling thread on the GPU's NUMA node, so that pinned host
    buffers allocated (first touched) inside it are local to the GPU's PCIe root and the
    copy-issuing thread is too. A no-op where NVML gives no answer."""

    def __init__(self, device_index=None):
        self.dev = torch.cuda.current_device() if device_index is None else device_index
        self.cpus = None
        self.saved = None

    def __enter__(self):
        self.cpus = cpus_near_gpu(self.dev)
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except OSError:
                self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved:
            os.sched_setaffinity(0, self.saved)
        return False


class HostStepPipeline:
    def __init__(self, ncells, nt, problem_type=7, chunk_nt=640, nslots=3,
                 reconstructors=("qlt", "caas"), rank=0, nranks=1, p2p=False):
        """With nranks > 1 (one process per GPU, torch.distributed initialised) this rank
        holds cells [rank*ncells/nranks, (rank+1)*ncells/nranks) of every tracer and each
        chunk's run() all-gathers the block roots (subtree partition)."""
        ncells_global = ncells
        ncells = ncells//nranks
        self.ncells, self.nt = ncells, nt
        self.chunk_nt = min(chunk_nt, nt)
        self.nchunks = (nt + self.chunk_nt - 1)//self.chunk_nt
        self.kinds = tuple(reconstructors)
        self.slots = []
        for _ in range(min(nslots, self.nchunks)):
            s = {"stream": torch.cuda.Stream()}
            with torch.cuda.stream(s["stream"]):
                s["in"] = [torch.empty((self.chunk_nt, ncells), dtype=torch.float64,
                                       device="cuda") for _ in range(4)]
                s["rhom"] = torch.empty(ncells, dtype=torch.float64, device="cuda")
                s["out"] = {k: torch.empty((self.chunk_nt, ncells), dtype=torch.float64,
                                           device="cuda") for k in self.kinds}
                s["cdr"] = {}
                for k in self.kinds:
                    if k == "qlt":
                        c = cb.QLT(ncells_global, rank=rank, nranks=nranks)
                    else:
                        c = cb.CAAS(ncells, cell0=rank*ncells, ncells_global=ncells_global,
                                    rank=rank, nranks=nranks)
                    for _t in range(self.chunk_nt):
                        c.declare_tracer(problem_type)
                    c.end_tracer_declarations()
                    if nranks > 1:
                        c.enable_distributed(nranks)
                    c.finish_setup()   # binds the slot's stream
                    if nranks > 1 and p2p:
                        c.enable_p2p(nranks)
                    s["cdr"][k] = c
                # Zero-copy: the CDRs read the slot's input arrays in place (no set_Qm /
                # get_Qm kernels); CAAS updates Qm in place, so it runs after QLT.
                self.bound = ncells % 2 == 0 and self.chunk_nt == nt or \
                    (ncells % 2 == 0 and nt % self.chunk_nt == 0)
                if self.bound:
                    lo, q, hi, prev = s["in"]
                    for k in self.kinds:
                        if k == "qlt":
                            s["cdr"][k].bind_arrays(q, lo, hi, prev, out=s["out"][k])
                        else:
                            s["cdr"][k].bind_arrays(q, lo, hi, prev)
            self.slots.append(s)
        torch.cuda.synchronize()
        self.h2d_bytes_per_step = 8*(4*nt*ncells + self.nchunks*ncells)
        self.d2h_bytes_per_step = 8*len(self.kinds)*nt*ncells
        self.launches_per_step = 0

    def step(self, rhom_h, qm_min_h, qm_h, qm_max_h, qm_prev_h, out_h):
        """Inputs: pinned host tensors [nt, ncells] (rhom_h [ncells]); out_h: dict
        kind -> pinned host tensor [nt, ncells]. Returns after all copies landed."""
        launches = 0
        for ci in range(self.nchunks):
            s = self.slots[ci % len(self.slots)]
            t0 = ci*self.chunk_nt
            n = min(self.chunk_nt, self.nt - t0)
            with torch.cuda.stream(s["stream"]):
                s["rhom"].copy_(rhom_h, non_blocking=True)
                for dst, src in zip(s["in"], (qm_min_h, qm_h, qm_max_h, qm_prev_h)):
                    dst[:n].copy_(src[t0:t0+n], non_blocking=True)
                lo, q, hi, prev = s["in"]
                for k in sorted(self.kinds, key=lambda x: x == "caas"):   # caas last (in place)
                    c = s["cdr"][k]
                    c.set_rhom(s["rhom"])
                    if self.bound:
                        c.run()
                        res = s["out"][k] if k == "qlt" else q
                        launches += c.last_run_launches()
                    else:
                        c.set_Qm(q, lo, hi, prev)
                        c.run()
                        c.get_Qm(out=s["out"][k])
                        res = s["out"][k]
                        launches += c.last_run_launches() + 2
                    out_h[k][t0:t0+n].copy_(res[:n], non_blocking=True)
        for s in self.slots:
            s["stream"].synchronize()
        self.launches_per_step = launches
