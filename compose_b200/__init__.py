"""compose_b200 -- B200-native CEDR property preservation (QLT and CAAS).

Python host-side mirror of the reference's `cedr::CDR` interface
(cedr/cedr_cdr.hpp:16-112) over the C ABI of include/cedr_b200.h. PyTorch is
used only for device memory and streams; every kernel is in
compose_b200/libcedr_b200.so (hand-written CUDA, sm_100a). There is no CPU
fallback: constructing a CDR without the library or without a GPU raises.

    qlt = QLT(ncells)                       # tree::make_tree_over_1d_mesh
    for t in range(nt): qlt.declare_tracer(CONSERVE | SHAPEPRESERVE | CONSISTENT)
    qlt.end_tracer_declarations(); qlt.finish_setup()
    qlt.set_rhom(rhom); qlt.set_Qm(qm, qm_min, qm_max, qm_prev)   # cuda float64
    qlt.run(); out = qlt.get_Qm()
"""
import ctypes as C
import os

CONSERVE, SHAPEPRESERVE, CONSISTENT, NONNEGATIVE = 1, 2, 4, 8
CAAS_SUM_TREE, CAAS_SUM_SEQUENTIAL, CAAS_SUM_USER = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
# CEDR_B200_LIB: a differently built copy of the library (kernel A/B experiments).
LIB_PATH = os.environ.get("CEDR_B200_LIB") or os.path.join(_HERE, "libcedr_b200.so")

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_int64)
_vp = C.c_void_p

ALLGATHER_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, _vp, C.c_size_t, _vp)
USER_REDUCER_FN = C.CFUNCTYPE(C.c_int, _vp, _vp, _vp, C.c_int, C.c_int, _vp)


class DeviceOp(C.Structure):
    """cedr_b200_device_op (include/cedr_b200.h)."""
    _fields_ = [("in_", _vp), ("out", _vp), ("ld", C.c_int64), ("trcr_row", _vp),
                ("trcr_prob", _vp), ("ntracers", C.c_int), ("nlclcells", C.c_int),
                ("is_caas", C.c_int), ("reserved", C.c_int)]


# Every symbol include/cedr_b200.h declares: (name, restype, argtypes).
_H = _vp  # opaque cedr_b200_cdr*
SYMBOLS = [
    ("cedr_b200_last_error", C.c_char_p, []),
    ("cedr_b200_version", C.c_int, []),
    ("cedr_b200_device_available", C.c_int, []),
    ("cedr_b200_qlt_create", C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, _ip, _lp,
                                       _ip, C.c_int, C.c_int, C.c_int]),
    ("cedr_b200_qlt_create_1d", C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int]),
    ("cedr_b200_caas_create", C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int64,
                                        C.c_int64, C.c_int, C.c_int]),
    ("cedr_b200_caas_set_user_reducer", C.c_int, [_H, USER_REDUCER_FN, _vp, C.c_int]),
    ("cedr_b200_bfb_create", C.c_int, [C.POINTER(_H), C.c_int, C.c_int, C.c_int, _ip, _lp,
                                       _ip, C.c_int, C.c_int, C.c_int, C.c_int]),
    ("cedr_b200_bfb_allreduce", C.c_int, [_H, _vp, _vp, C.c_int, C.c_int]),
    ("cedr_b200_destroy", C.c_int, [_H]),
    ("cedr_b200_declare_tracer", C.c_int, [_H, C.c_int, C.c_int]),
    ("cedr_b200_end_tracer_declarations", C.c_int, [_H]),
    ("cedr_b200_get_buffers_sizes", C.c_int, [_H, C.POINTER(C.c_size_t),
                                              C.POINTER(C.c_size_t)]),
    ("cedr_b200_set_buffers", C.c_int, [_H, _vp, _vp]),
    ("cedr_b200_finish_setup", C.c_int, [_H]),
    ("cedr_b200_get_problem_type", C.c_int, [_H, C.c_int, _ip]),
    ("cedr_b200_get_num_tracers", C.c_int, [_H, _ip]),
    ("cedr_b200_run", C.c_int, [_H]),
    ("cedr_b200_print", C.c_int, [_H, C.c_char_p, C.c_size_t]),
    ("cedr_b200_nlclcells", C.c_int, [_H, _ip]),
    ("cedr_b200_get_owned_glblcells", C.c_int, [_H, _lp]),
    ("cedr_b200_gci2lci", C.c_int, [_H, C.c_int64, _ip]),
    ("cedr_b200_get_device_op", C.c_int, [_H, C.POINTER(DeviceOp)]),
    ("cedr_b200_set_rhom_bulk", C.c_int, [_H, _vp]),
    ("cedr_b200_set_Qm_bulk", C.c_int, [_H, C.c_int, C.c_int, C.c_int64, _vp, _vp, _vp,
                                        _vp]),
    ("cedr_b200_get_Qm_bulk", C.c_int, [_H, C.c_int, C.c_int, C.c_int64, _vp]),
    ("cedr_b200_bind_arrays", C.c_int, [_H, C.c_int64, _vp, _vp, _vp, _vp, _vp]),
    ("cedr_b200_local_solve", C.c_int, [C.c_int, C.c_int, C.c_int] + [_vp]*8 +
     [C.c_int64, C.c_int64, C.c_int, C.c_int, _vp]),
    ("cedr_b200_transport1d_cycle", C.c_int, [_H, C.c_int, _dp, _dp, C.c_int,
                                              C.POINTER(C.c_float)]),
    ("cedr_b200_set_stream", C.c_int, [_H, _vp]),
    ("cedr_b200_synchronize", C.c_int, [_H]),
    ("cedr_b200_set_allgather", C.c_int, [_H, ALLGATHER_FN, _vp]),
    ("cedr_b200_get_exchange_count", C.c_int, [_H, C.POINTER(C.c_size_t)]),
    ("cedr_b200_set_exchange_buffers", C.c_int, [_H, _vp, _vp]),
    ("cedr_b200_get_exchange_buffers", C.c_int, [_H, C.POINTER(_vp), C.POINTER(_vp)]),
    ("cedr_b200_run_phase", C.c_int, [_H, C.c_int]),
    ("cedr_b200_partition_probe", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_int, _ip, _ip, _ip, _ip, _ip, _ip, _ip]),
    ("cedr_b200_p2p_get_handle", C.c_int, [_H, _vp]),
    ("cedr_b200_p2p_set_peer", C.c_int, [_H, C.c_int, _vp]),
    ("cedr_b200_p2p_enable", C.c_int, [_H, C.c_int]),
    ("cedr_b200_last_run_launches", C.c_int, [_H, _ip]),
    ("cedr_b200_set_fast_path", C.c_int, [_H, C.c_int]),
    ("cedr_b200_uses_fast_path", C.c_int, [_H, _ip]),
    ("cedr_b200_debug_phase_clocks", C.c_int, [_H, C.POINTER(C.c_ulonglong)]),
    ("cedr_b200_set_ring", C.c_int, [_H, C.c_int]),
    ("cedr_b200_uses_ring", C.c_int, [_H, _ip]),
    ("cedr_b200_set_cluster_caas", C.c_int, [_H, C.c_int]),
    ("cedr_b200_uses_cluster_caas", C.c_int, [_H, _ip]),
    ("cedr_b200_set_graph", C.c_int, [_H, C.c_int]),
    ("cedr_b200_uses_graph", C.c_int, [_H, _ip]),
    ("cedr_b200_ring_info", C.c_int, [_H, _ip]),
    ("cedr_b200_ring_trace", C.c_int, [_H, C.POINTER(C.c_ulonglong), C.c_size_t,
                                       C.POINTER(C.c_size_t)]),
    ("cedr_b200_set_profiling", C.c_int, [_H, C.c_int]),
    ("cedr_b200_get_launch_times", C.c_int, [_H, C.c_int, C.POINTER(C.c_float), _ip, _ip,
                                             _ip]),
    ("cedr_b200_plan_info", C.c_int, [_H, _ip, _ip, _ip, _ip]),
    ("cedr_b200_set_max_block_leaves", C.c_int, [_H, C.c_int]),
    ("cedr_b200_plan_probe", C.c_int, [C.c_int, C.c_int, C.c_int, _ip, _lp, C.c_int, _lp,
                                       _ip, _ip, _ip, _ip, _lp]),
    ("cedr_b200_make_1d_tree", C.c_int, [C.c_int, C.c_int, _ip, _lp]),
    ("cedr_b200_merge_partial_trees", C.c_int, [C.c_int, _ip, _ip, _ip, _lp, _ip, C.c_int,
                                                _ip, _ip, _lp, _ip]),
    ("cedr_b200_allgather_host", C.c_int, [ALLGATHER_FN, _vp, C.c_int, _vp, _vp,
                                           C.c_size_t]),
    ("cedr_b200_fill_headline", C.c_int, [C.c_int, C.c_int, C.c_int64, C.c_int, _vp, _vp,
                                          _vp, _vp, _vp, _vp]),
    ("cedr_b200_fill_headline_range", C.c_int, [C.c_int, C.c_int, C.c_int, C.c_int,
                                                C.c_int64, C.c_int, _vp, _vp, _vp, _vp, _vp,
                                                _vp]),
]

_lib = None


def load_library(path=None):
    """dlopen the CUDA library and bind every declared symbol. Raises if absent."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    p = path or LIB_PATH
    if not os.path.exists(p):
        raise RuntimeError(
            "compose_b200: %s is missing. Build it with `python -m compose_b200.build` "
            "(nvcc, sm_100a). There is no CPU fallback." % p)
    lib = C.CDLL(p)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)  # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    if path is None:
        _lib = lib
    return lib


class CedrError(RuntimeError):
    """A failure reported by the C ABI (the reference throws std::logic_error)."""


def _check(rc):
    if rc:
        raise CedrError(load_library().cedr_b200_last_error().decode())


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _row_stride(x):
    """Leading dimension of a [nt, n] SoA tensor whose rows are unit-stride."""
    assert x.dim() == 2 and (x.shape[1] == 1 or x.stride(1) == 1)
    if x.shape[0] == 1 or x.shape[1] == 1:
        # strides of size-1 dimensions are arbitrary
        return max(x.shape[1], x.stride(0) if x.shape[0] > 1 and x.shape[1] > 1 else 0)
    return x.stride(0)


def _as_i32(a):
    import numpy as np
    return np.ascontiguousarray(a, dtype=np.int32)


class CDR:
    """Host methods of cedr::CDR (cedr_cdr.hpp:47-108), plus bulk DeviceOp forms."""

    def __init__(self):
        self._lib = load_library()
        self._h = _H()
        self._in = self._out = None      # torch tensors backing the buffers
        self._cb = None

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self._lib.cedr_b200_destroy(self._h)
                self._h = None
        except Exception:
            pass

    # -- setup
    def declare_tracer(self, problem_type, rhomidx=0):
        _check(self._lib.cedr_b200_declare_tracer(self._h, int(problem_type), int(rhomidx)))

    def end_tracer_declarations(self):
        _check(self._lib.cedr_b200_end_tracer_declarations(self._h))

    def get_buffers_sizes(self):
        b1, b2 = C.c_size_t(0), C.c_size_t(0)
        _check(self._lib.cedr_b200_get_buffers_sizes(self._h, C.byref(b1), C.byref(b2)))
        return b1.value, b2.value

    def set_buffers(self, buf1, buf2):
        """buf1/buf2: cuda float64 tensors of at least get_buffers_sizes() elements."""
        self._in, self._out = buf1, buf2
        _check(self._lib.cedr_b200_set_buffers(self._h, _ptr(buf1), _ptr(buf2)))

    def set_max_block_leaves(self, n):
        _check(self._lib.cedr_b200_set_max_block_leaves(self._h, int(n)))

    def finish_setup(self):
        self.use_current_stream()
        _check(self._lib.cedr_b200_finish_setup(self._h))

    def get_problem_type(self, tracer_idx):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_get_problem_type(self._h, int(tracer_idx), C.byref(v)))
        return v.value

    def get_num_tracers(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_get_num_tracers(self._h, C.byref(v)))
        return v.value

    def nlclcells(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_nlclcells(self._h, C.byref(v)))
        return v.value

    def get_owned_glblcells(self):
        import numpy as np
        out = np.empty(self.nlclcells(), np.int64)
        _check(self._lib.cedr_b200_get_owned_glblcells(self._h, out.ctypes.data_as(_lp)))
        return out

    def gci2lci(self, gci):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_gci2lci(self._h, int(gci), C.byref(v)))
        return v.value

    def print(self):
        buf = C.create_string_buffer(4096)
        _check(self._lib.cedr_b200_print(self._h, buf, 4096))
        return buf.value.decode()

    def plan_info(self):
        a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
        _check(self._lib.cedr_b200_plan_info(self._h, C.byref(a), C.byref(b), C.byref(c),
                                             C.byref(d)))
        return {"ntiers": a.value, "nblocks0": b.value, "max_block_leaves": c.value,
                "nlevels_ref": d.value}

    def get_device_op(self):
        op = DeviceOp()
        _check(self._lib.cedr_b200_get_device_op(self._h, C.byref(op)))
        return op

    # -- streams
    def use_current_stream(self):
        _check(self._lib.cedr_b200_set_stream(self._h, _stream_ptr()))

    def synchronize(self):
        _check(self._lib.cedr_b200_synchronize(self._h))

    # -- per step
    def set_rhom(self, rhom):
        """rhom: cuda float64 [nlclcells], indexed by local cell index."""
        assert rhom.is_cuda and rhom.dtype.is_floating_point and rhom.element_size() == 8
        _check(self._lib.cedr_b200_set_rhom_bulk(self._h, _ptr(rhom)))

    def set_Qm(self, qm, qm_min, qm_max, qm_prev=None, t0=0):
        """SoA cuda float64 [nt, lda] arrays, cell (lci) fastest."""
        nt, lda = qm.shape[0], _row_stride(qm)
        for x in (qm, qm_min, qm_max, qm_prev):
            assert x is None or (x.is_cuda and x.element_size() == 8 and
                                 x.shape == qm.shape and _row_stride(x) == lda)
        _check(self._lib.cedr_b200_set_Qm_bulk(self._h, int(t0), int(nt), int(lda), _ptr(qm),
                                               _ptr(qm_min), _ptr(qm_max), _ptr(qm_prev)))

    def bind_arrays(self, qm, qm_min, qm_max, qm_prev=None, out=None):
        """Zero-copy set_Qm/get_Qm: run() reads these SoA cuda float64 [nt, lda] arrays in
        place and writes QLT's results to `out` (CAAS updates qm in place). bind_arrays(None,
        None, None) unbinds. The tensors must outlive the binding."""
        if qm is None:
            self._bound = None
            _check(self._lib.cedr_b200_bind_arrays(self._h, 0, None, None, None, None, None))
            return
        lda = _row_stride(qm)
        for x in (qm, qm_min, qm_max, qm_prev, out):
            assert x is None or (x.is_cuda and x.element_size() == 8 and
                                 x.shape == qm.shape and _row_stride(x) == lda)
        self._bound = (qm, qm_min, qm_max, qm_prev, out)
        _check(self._lib.cedr_b200_bind_arrays(self._h, int(lda), _ptr(qm_min), _ptr(qm),
                                               _ptr(qm_max), _ptr(qm_prev), _ptr(out)))

    def transport1d_cycle(self, nsteps, y0, use_graph=True):
        """Problem1D::cycle (cedr_test_1d_transport.cpp:231-254) on the device for tracer 0.
        y0: numpy [ncells + 1]. Returns (yf, device microseconds per step). use_graph:
        False = three launches per step, True = replayed from a CUDA graph, 2 = the whole
        cycle in one launch (one block of <= 256 cells, one tracer)."""
        import numpy as np
        y0 = np.ascontiguousarray(y0, dtype=np.float64)
        yf = np.empty_like(y0)
        ms = C.c_float(0)
        _check(self._lib.cedr_b200_transport1d_cycle(
            self._h, int(nsteps), y0.ctypes.data_as(_dp), yf.ctypes.data_as(_dp),
            2 if use_graph == 2 and use_graph is not True else int(bool(use_graph)),
            C.byref(ms)))
        return yf, 1e3*ms.value

    def run(self):
        _check(self._lib.cedr_b200_run(self._h))

    # -- multi-rank (subtree partition; one process per GPU)
    def exchange_count(self):
        v = C.c_size_t(0)
        _check(self._lib.cedr_b200_get_exchange_count(self._h, C.byref(v)))
        return v.value

    def enable_distributed(self, nranks, group=None):
        """Call between end_tracer_declarations and finish_setup on every rank: allocates
        the exchange message buffers as torch tensors and wires run()'s one all-gather to
        torch.distributed (NCCL over NVLink on a GPU box)."""
        import torch
        import torch.distributed as dist
        n = self.exchange_count()
        self._xsend = torch.zeros(max(n, 1), dtype=torch.float64, device="cuda")
        self._xrecv = torch.zeros(max(n, 1)*nranks, dtype=torch.float64, device="cuda")
        _check(self._lib.cedr_b200_set_exchange_buffers(self._h, _ptr(self._xsend),
                                                        _ptr(self._xrecv)))

        def gather(ctx, send, recv, count, stream):
            # Enqueue the collective on the CDR's own stream (the pack / unpack kernels
            # either side of it run there), whatever torch's current stream is.
            try:
                sp = int(stream or 0)
                if sp == torch.cuda.current_stream().cuda_stream:
                    dist.all_gather_into_tensor(self._xrecv, self._xsend, group=group)
                else:
                    with torch.cuda.stream(torch.cuda.ExternalStream(sp)):
                        dist.all_gather_into_tensor(self._xrecv, self._xsend, group=group)
                return 0
            except Exception:   # surfaced as a CedrError by run()
                import traceback
                traceback.print_exc()
                return 1
        self._cb = ALLGATHER_FN(gather)
        _check(self._lib.cedr_b200_set_allgather(self._h, self._cb, None))

    def enable_p2p(self, nranks, group=None):
        """After finish_setup on every rank: swap CUDA IPC handles through
        torch.distributed and switch run()'s exchange to direct stores into the peers'
        buffers over NVLink (no NCCL call inside run())."""
        import torch
        import torch.distributed as dist
        buf = C.create_string_buffer(64)
        ok = 1
        try:
            _check(self._lib.cedr_b200_p2p_get_handle(self._h, buf))
        except CedrError:
            ok = 0
        handles = [None]*nranks
        dist.all_gather_object(handles, bytes(buf.raw) if ok else None, group=group)
        if ok and all(h is not None for h in handles):
            try:
                for r, h in enumerate(handles):
                    _check(self._lib.cedr_b200_p2p_set_peer(self._h, r,
                                                            C.create_string_buffer(h, 64)))
            except CedrError:
                ok = 0
        else:
            ok = 0
        # Every rank must take the same path: peer mapping can fail on one (no IPC in the
        # container, no peer access); then all keep the all-gather callback.
        t = torch.tensor([ok], dtype=torch.int32, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        ok = int(t.item())
        _check(self._lib.cedr_b200_p2p_enable(self._h, ok))
        return bool(ok)

    def run_phase(self, phase):
        _check(self._lib.cedr_b200_run_phase(self._h, int(phase)))

    def exchange_buffers(self, nranks):
        """(send, recv) as cuda float64 tensors viewing the CDR's message buffers; only
        for emulating several ranks in one process (tests)."""
        return self._xsend, self._xrecv

    def use_tensor_exchange_buffers(self, nranks):
        import torch
        n = self.exchange_count()
        self._xsend = torch.zeros(max(n, 1), dtype=torch.float64, device="cuda")
        self._xrecv = torch.zeros(max(n, 1)*nranks, dtype=torch.float64, device="cuda")
        _check(self._lib.cedr_b200_set_exchange_buffers(self._h, _ptr(self._xsend),
                                                        _ptr(self._xrecv)))

    def get_Qm(self, out=None, t0=0, nt=None):
        import torch
        nt = self.get_num_tracers() - t0 if nt is None else nt
        if out is None:
            out = torch.empty((nt, self.nlclcells()), dtype=torch.float64, device="cuda")
        lda = _row_stride(out)
        _check(self._lib.cedr_b200_get_Qm_bulk(self._h, int(t0), int(nt), int(lda),
                                               _ptr(out)))
        return out

    def set_fast_path(self, on=True):
        _check(self._lib.cedr_b200_set_fast_path(self._h, int(bool(on))))

    def uses_fast_path(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_uses_fast_path(self._h, C.byref(v)))
        return bool(v.value)

    def set_ring(self, on=True):
        _check(self._lib.cedr_b200_set_ring(self._h, int(bool(on))))

    def set_graph(self, mode):
        """run() as a replayed CUDA graph: -1 auto (multi-rank peer-to-peer runs), 0 never,
        1 whenever run() is pure stream work (cedr_b200_set_graph)."""
        _check(self._lib.cedr_b200_set_graph(self._h, int(mode)))

    def uses_graph(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_uses_graph(self._h, C.byref(v)))
        return bool(v.value)

    def uses_ring(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_uses_ring(self._h, C.byref(v)))
        return bool(v.value)

    def set_cluster_caas(self, mode):
        """Opt-in: CAAS::run as one cluster kernel (a tracer on chip, one pass over HBM):
        1 on where it applies, 0 off (cedr_b200_set_cluster_caas)."""
        _check(self._lib.cedr_b200_set_cluster_caas(self._h, int(mode)))

    def uses_cluster_caas(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_uses_cluster_caas(self._h, C.byref(v)))
        return bool(v.value)

    def ring_info(self):
        v = (C.c_int*8)()
        _check(self._lib.cedr_b200_ring_info(self._h, v))
        keys = ("grid", "S", "npn", "TB", "nslots", "np", "sw", "smem")
        d = dict(zip(keys, list(v)))
        d["nuslots"], d["ndslots"] = divmod(d.pop("nslots"), 100)
        return d

    def ring_trace(self):
        """Debug (CEDR_B200_RING_TRACE=1): numpy uint64 stamps of the last ring launch."""
        import numpy as np
        n = C.c_size_t(0)
        _check(self._lib.cedr_b200_ring_trace(self._h, None, 0, C.byref(n)))
        out = np.zeros(n.value, np.uint64)
        if n.value:
            _check(self._lib.cedr_b200_ring_trace(
                self._h, out.ctypes.data_as(C.POINTER(C.c_ulonglong)), n.value, C.byref(n)))
        return out

    def debug_phase_clocks(self):
        out = (C.c_ulonglong*16)()
        _check(self._lib.cedr_b200_debug_phase_clocks(self._h, out))
        return list(out)

    def set_profiling(self, on=True):
        _check(self._lib.cedr_b200_set_profiling(self._h, int(bool(on))))

    def launch_times(self):
        """[(tag, tier, ms)] of the last run() (profiling must be on)."""
        cap = 4096
        ms = (C.c_float*cap)()
        tags, tiers, n = (C.c_int*cap)(), (C.c_int*cap)(), C.c_int(0)
        _check(self._lib.cedr_b200_get_launch_times(self._h, cap, ms, tags, tiers,
                                                    C.byref(n)))
        names = ["rhom", "up", "top", "down", "caas_adjust", "exchange", "fused", "mid"]
        return [(names[tags[i]], tiers[i], ms[i]) for i in range(n.value)]

    def last_run_launches(self):
        v = C.c_int(0)
        _check(self._lib.cedr_b200_last_run_launches(self._h, C.byref(v)))
        return v.value


class QLT(CDR):
    """cedr::qlt::QLT (cedr_qlt.hpp:26-219).

    tree=None builds tree::make_tree_over_1d_mesh(ncells, imbalanced); otherwise
    `tree` is (kids[2*nnodes], cellidx[nnodes], root) flat arrays of the caller's
    tree::Node graph.
    """

    def __init__(self, ncells, tree=None, imbalanced=False,
                 prefer_numerical_mass_conservation_to_numerical_bounds=False,
                 rank=0, nranks=1, node_rank=None):
        super().__init__()
        import numpy as np
        prefer = int(bool(prefer_numerical_mass_conservation_to_numerical_bounds))
        if tree is None:
            _check(self._lib.cedr_b200_qlt_create_1d(C.byref(self._h), int(ncells),
                                                     int(bool(imbalanced)), prefer,
                                                     int(rank), int(nranks)))
        else:
            kids, cellidx, root = tree
            kids = _as_i32(kids).reshape(-1)
            cellidx = np.ascontiguousarray(cellidx, dtype=np.int64)
            nr = None if node_rank is None else _as_i32(node_rank)
            _check(self._lib.cedr_b200_qlt_create(
                C.byref(self._h), int(ncells), int(cellidx.size), int(root),
                kids.ctypes.data_as(_ip), cellidx.ctypes.data_as(_lp),
                None if nr is None else nr.ctypes.data_as(_ip), prefer, int(rank),
                int(nranks)))


class BfbTreeAllReducer(CDR):
    """cedr::BfbTreeAllReducer (cedr_bfb_tree_allreduce.hpp:15-55), device-resident."""

    def __init__(self, nleaf, nfield, tree=None, max_block_leaves=0, rank=0, nranks=1,
                 node_rank=None):
        super().__init__()
        import numpy as np
        self.nfield = nfield
        if tree is None:
            args = (0, 0, None, None, None)
        else:
            kids, cellidx, root = tree
            kids = _as_i32(kids).reshape(-1)
            cellidx = np.ascontiguousarray(cellidx, dtype=np.int64)
            nr = None if node_rank is None else _as_i32(node_rank)
            self._keep = (kids, cellidx, nr)
            args = (int(cellidx.size), int(root), kids.ctypes.data_as(_ip),
                    cellidx.ctypes.data_as(_lp), None if nr is None else nr.ctypes.data_as(_ip))
        _check(self._lib.cedr_b200_bfb_create(C.byref(self._h), int(nleaf), *args, int(nfield),
                                              int(max_block_leaves), int(rank), int(nranks)))
        self.use_current_stream()

    def allreduce(self, send, recv=None, transpose=False, phase=-1):
        import torch
        if recv is None:
            recv = torch.empty(self.nfield, dtype=torch.float64, device="cuda")
        _check(self._lib.cedr_b200_bfb_allreduce(self._h, _ptr(send), _ptr(recv),
                                                 int(bool(transpose)), int(phase)))
        return recv


class CAAS(CDR):
    """cedr::caas::CAAS (cedr_caas.hpp:15-118).

    user_reducer: the reference's UserAllReducer (cedr_caas.hpp:27-49) as a callable
    ``f(send, recv, nlocal, nfld)`` on cuda float64 tensors -- send viewed as
    [nfld, nlocal] (nlocal fastest), recv [nfld] to fill with the sums; n_accum is its
    n_accum_in_place(). With a reducer this rank's cells may be any set."""

    def __init__(self, nlclcells, sum_mode=CAAS_SUM_TREE, cell0=0, ncells_global=None,
                 rank=0, nranks=1, user_reducer=None, n_accum=1):
        super().__init__()
        ng = nlclcells if ncells_global is None else ncells_global
        if user_reducer is not None:
            sum_mode = CAAS_SUM_USER
        _check(self._lib.cedr_b200_caas_create(C.byref(self._h), int(nlclcells),
                                               int(sum_mode), int(cell0), int(ng),
                                               int(rank), int(nranks)))
        if user_reducer is not None:
            import torch

            def tramp(ctx, send, recv, nlocal, nfld, stream):
                try:
                    s = _wrap_device(send, nlocal*nfld).view(nfld, nlocal)
                    r = _wrap_device(recv, nfld)
                    user_reducer(s, r, nlocal, nfld)
                    torch.cuda.synchronize()
                    return 0
                except Exception:   # surfaced as a CedrError by run()
                    import traceback
                    traceback.print_exc()
                    return 1
            self._ucb = USER_REDUCER_FN(tramp)
            _check(self._lib.cedr_b200_caas_set_user_reducer(self._h, self._ucb, None,
                                                             int(n_accum)))


def allreduce_sum_reducer(group=None):
    """The reference's default global reduction (cedr_caas.cpp:203-209: one partial sum per
    field and rank, accumulated cell after cell, then MPI_Allreduce(SUM)) as a UserAllReducer
    over torch.distributed. Use with n_accum = nlclcells, so that each rank contributes its
    single sequential partial; this rank's cells may then be ANY set (cedr_caas.cpp:37-48
    takes only nlclcells). The order in which the ranks' partials are added is the
    collective's, as it is MPI's in the reference (two ranks: exact either way)."""
    def reducer(send, recv, nlocal, nfld):
        import torch.distributed as dist
        if nlocal != 1:
            raise ValueError("allreduce_sum_reducer needs n_accum == nlclcells (one partial "
                             "per rank); got %d partials" % nlocal)
        recv.copy_(send[:, 0])
        dist.all_reduce(recv, op=dist.ReduceOp.SUM, group=group)
    return reducer


def _wrap_device(ptr, n):
    """A cuda float64 tensor over n doubles at device address `ptr` (no copy)."""
    import torch

    class _Arr:
        pass
    a = _Arr()
    a.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False),
                                  "version": 2}
    return torch.as_tensor(a, device="cuda")


def fill_headline(ncells, nt, config_id, lda=None, cell0=0, nlclcells=None):
    """Synthetic SURVEY 8(d) workload generated on the device, optionally only cells
    [cell0, cell0 + nlclcells). Returns cuda tensors (rhom[nlcl], qm_min, qm, qm_max,
    qm_prev: [nt, lda])."""
    import torch
    lib = load_library()
    nl = ncells - cell0 if nlclcells is None else nlclcells
    lda = nl if lda is None else lda
    rhom = torch.empty(nl, dtype=torch.float64, device="cuda")
    arrs = [torch.empty((nt, lda), dtype=torch.float64, device="cuda") for _ in range(4)]
    _check(lib.cedr_b200_fill_headline_range(int(ncells), int(cell0), int(nl), int(nt),
                                             int(lda), int(config_id), _ptr(rhom),
                                             _ptr(arrs[0]), _ptr(arrs[1]), _ptr(arrs[2]),
                                             _ptr(arrs[3]), _stream_ptr()))
    return (rhom,) + tuple(arrs)


LOCAL_QP, LOCAL_CAAS, LOCAL_NONNEG_LS, LOCAL_NONNEG_CAAS, LOCAL_QP_2D = range(5)


def local_solve(method, b, y, xlo=None, xhi=None, w=None, a=None, max_its=0, clip=True):
    """cedr::local solvers (cedr_local.hpp:23-58) for a batch of elements: SoA cuda float64
    arrays [n, nprob] (element index fastest), b [nprob]. Returns (x [n, nprob], info)."""
    import torch
    n, nprob = y.shape
    for t in (y, xlo, xhi, w, a):
        assert t is None or (t.is_cuda and t.is_contiguous() and t.shape == y.shape)
    x = torch.zeros_like(y)
    info = torch.zeros(nprob, dtype=torch.int32, device="cuda")
    _check(load_library().cedr_b200_local_solve(
        int(method), int(nprob), int(n), _ptr(w), _ptr(a), _ptr(b), _ptr(xlo), _ptr(xhi),
        _ptr(y), _ptr(x), _ptr(info), int(nprob), 1, int(max_its), int(bool(clip)),
        _stream_ptr()))
    return x, info


def make_1d_tree(ncells, imbalanced=False):
    """tree::make_tree_over_1d_mesh as flat arrays: (kids, cellidx, root)."""
    import numpy as np
    lib = load_library()
    nn = 2*ncells - 1
    kids = np.empty(2*nn, np.int32)
    cellidx = np.empty(nn, np.int64)
    _check(lib.cedr_b200_make_1d_tree(int(ncells), int(bool(imbalanced)),
                                      kids.ctypes.data_as(_ip), cellidx.ctypes.data_as(_lp)))
    return kids, cellidx, 0


def merge_partial_trees(parts):
    """Union of the ranks' partial trees (tree::Node::level, cedr_tree_caller.hpp:20-22).

    `parts[p]` = (kids, cellidx, root, node_rank) of rank p: the global tree with the
    subtrees that hold none of rank p's cells cut down to kid-less stubs. Returns
    ((kids, cellidx, 0), node_rank) of the whole tree in pre-order, ready for QLT(...).
    """
    import numpy as np
    lib = load_library()
    nn = np.array([np.asarray(t[1]).size for t in parts], np.int32)
    roots = np.array([int(t[2]) for t in parts], np.int32)
    kids = np.concatenate([_as_i32(t[0]).reshape(-1) for t in parts])
    cellidx = np.concatenate([np.asarray(t[1], np.int64).reshape(-1) for t in parts])
    rank = np.concatenate([_as_i32(t[3]).reshape(-1) for t in parts])
    cap = int(nn.sum())
    ok, oc, orank = np.empty(2*cap, np.int32), np.empty(cap, np.int64), np.empty(cap, np.int32)
    n = C.c_int(0)
    _check(lib.cedr_b200_merge_partial_trees(
        len(parts), nn.ctypes.data_as(_ip), roots.ctypes.data_as(_ip),
        kids.ctypes.data_as(_ip), cellidx.ctypes.data_as(_lp), rank.ctypes.data_as(_ip),
        cap, C.byref(n), ok.ctypes.data_as(_ip), oc.ctypes.data_as(_lp),
        orank.ctypes.data_as(_ip)))
    n = n.value
    return (ok[:2*n].copy(), oc[:n].copy(), 0), orank[:n].copy()


def assemble_partial_tree(tree, node_rank, group=None):
    """Every rank passes its own partial tree; all get the merged whole tree back.

    One setup-time all-gather of the parts over `group` (torch.distributed; any backend)
    stands in for what the reference's level schedule gets from MPI at run time
    (cedr_tree.cpp:71-76): the block plan needs the whole tree on every rank.
    """
    import torch.distributed as dist
    kids, cellidx, root = tree
    parts = [None]*dist.get_world_size(group)
    dist.all_gather_object(parts, (kids, cellidx, int(root), node_rank), group=group)
    return merge_partial_trees(parts)


def partition_probe(ncells, rank, nranks, max_block_leaves=1024, imbalanced=False):
    """Host-only: which tier-0 blocks of the 1-D mesh tree `rank` of `nranks` owns."""
    import numpy as np
    lib = load_library()
    cap = ncells
    g, l0, nl = (np.zeros(cap, np.int32) for _ in range(3))
    a, b, c, d = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int(0)
    _check(lib.cedr_b200_partition_probe(int(ncells), int(bool(imbalanced)),
                                         int(max_block_leaves), int(rank), int(nranks), cap,
                                         C.byref(a), C.byref(b), C.byref(c), C.byref(d),
                                         g.ctypes.data_as(_ip), l0.ctypes.data_as(_ip),
                                         nl.ctypes.data_as(_ip)))
    n = b.value
    return {"nlclcells": a.value, "nown": n, "nown_max": c.value, "nblocks": d.value,
            "gidx": g[:n].copy(), "leaf0": l0[:n].copy(), "nl": nl[:n].copy()}


def plan_probe(ncells, tree, max_block_leaves=1024):
    """Host-only: build the block plan for `tree` and run its self-check."""
    import numpy as np
    lib = load_library()
    kids, cellidx, root = tree
    kids = _as_i32(kids).reshape(-1)
    cellidx = np.ascontiguousarray(cellidx, dtype=np.int64)
    lci2gci = np.empty(ncells, np.int64)
    nb = np.zeros(64, np.int32)
    ntiers, nshapes, nlev, idsum = C.c_int(0), C.c_int(0), C.c_int(0), C.c_int64(0)
    _check(lib.cedr_b200_plan_probe(int(ncells), int(cellidx.size), int(root),
                                    kids.ctypes.data_as(_ip), cellidx.ctypes.data_as(_lp),
                                    int(max_block_leaves), lci2gci.ctypes.data_as(_lp),
                                    C.byref(ntiers), nb.ctypes.data_as(_ip),
                                    C.byref(nshapes), C.byref(nlev), C.byref(idsum)))
    return {"lci2gci": lci2gci, "ntiers": ntiers.value,
            "nblocks": nb[:ntiers.value].tolist(), "nshapes": nshapes.value,
            "nlevels_ref": nlev.value, "idsum": idsum.value}
