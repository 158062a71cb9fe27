"""ctypes binding of libnjode_b200.so (C ABI in include/njode_b200.h) and the per-batch staging
logic.  There is no CPU fallback: ``cuda_lib()`` raises when the CUDA library is missing and the
runner refuses non-CUDA devices.  (``Lib`` can be pointed at another build of the same ABI; the
test-suite uses that to run the host *simulation* of the kernel source, never the product.)"""
import contextlib
import ctypes as C
import os

import numpy as np
import torch

from . import schedule as _sched

MAX_LINEAR = 8
RECOMPUTE_AUTO_BYTES = 256 << 20      # "auto": keep the [S, B, H] history of the forward pass up to this size
# saved hidden activations of the ODE network (njode_plan_t.act_bytes: the segment backward reads them instead of
# recomputing two layers per Euler step).  Measured on B200 (profiles/r2w_*): a gain while the records are a few hundred
# MB (5 000 demo paths x 100 steps = 208 MB: step 2.38 -> 2.23 ms; 1 000 paths: 1.79 -> 1.63; 2x100 nets, 5 000 paths:
# 7.77 -> 7.53), a loss at 832 MB (20 000 paths: the forward pays 0.25 ms for the stores, the backward gains 0.08 ms), so
# they are kept up to this budget; NJODE_SAVE_ACTIVATIONS=0 turns them off
SAVE_ACTIVATIONS_MAX_BYTES = int(os.environ.get("NJODE_SAVE_ACTIVATIONS_MAX_MB", "512")) << 20
_HERE = os.path.dirname(os.path.abspath(__file__))
CUDA_LIB_PATH = os.path.join(_HERE, "libnjode_b200.so")


class MlpT(C.Structure):
    _fields_ = [("n_linear", C.c_int32), ("dims", C.c_int32 * (MAX_LINEAR + 1)),
                ("act", C.c_int32 * MAX_LINEAR), ("w_off", C.c_int64 * MAX_LINEAR),
                ("b_off", C.c_int64 * MAX_LINEAR)]


class ModelT(C.Structure):
    _fields_ = [("input_size", C.c_int32), ("hidden_size", C.c_int32), ("output_size", C.c_int32),
                ("masked", C.c_int32), ("input_current_t", C.c_int32), ("loss_kind", C.c_int32),
                ("residual", C.c_int32), ("training", C.c_int32), ("weight", C.c_float),
                ("dropout_p", C.c_float), ("dropout_seed", C.c_uint64), ("n_params", C.c_int64),
                ("use_rnn", C.c_int32), ("reserved1", C.c_int32), ("net", MlpT * 5)]


class BatchT(C.Structure):
    _fields_ = [("B", C.c_int32), ("N", C.c_int32), ("K", C.c_int32), ("S", C.c_int32),
                ("E", C.c_int32), ("batch_size_norm", C.c_int32), ("path_id_offset", C.c_int32),
                ("n_units", C.c_int32), ("unit_kind", C.c_int32), ("n_loss_units", C.c_int32),
                ("seg_n1", C.c_int32 * 2), ("seg_n2", C.c_int32 * 2), ("reserved0", C.c_int32),
                ("X", C.c_void_p), ("M", C.c_void_p), ("start_X", C.c_void_p), ("n_obs_ot", C.c_void_p),
                ("path_ptr", C.c_void_p), ("path_rows", C.c_void_p), ("row_jump", C.c_void_p),
                ("step_dt", C.c_void_p), ("step_t", C.c_void_p), ("jump_step", C.c_void_p),
                ("jump_tau", C.c_void_p), ("step_event", C.c_void_p), ("jump_event", C.c_void_p),
                ("unit_desc", C.c_void_p)]


class PlanT(C.Structure):
    _fields_ = [("tile_paths", C.c_int32), ("threads", C.c_int32), ("grid_fwd", C.c_int32),
                ("grid_bwd", C.c_int32), ("weights_in_smem", C.c_int32), ("grads_in_smem", C.c_int32),
                ("smem_fwd_bytes", C.c_int64), ("smem_bwd_bytes", C.c_int64),
                ("image_floats", C.c_int64), ("workspace_bytes", C.c_int64), ("recompute_bytes", C.c_int64), ("act_bytes", C.c_int64)]


class SavedT(C.Structure):
    _fields_ = [("h_hist", C.c_void_p), ("h_before", C.c_void_p), ("y_after", C.c_void_p), ("act_hist", C.c_void_p)]


class SdeT(C.Structure):
    _fields_ = [("model", C.c_int32), ("dimension", C.c_int32), ("nb_steps", C.c_int32),
                ("return_vol", C.c_int32), ("drift", C.c_double), ("volatility", C.c_double),
                ("mean", C.c_double), ("speed", C.c_double), ("correlation", C.c_double),
                ("v0", C.c_double), ("maturity", C.c_double), ("sine_coeff", C.c_double),
                ("obs_perc", C.c_double), ("t0", C.c_double), ("seed", C.c_uint64)]


ACT_CODES = {"tanh": 1, "relu": 2}
LOSS_CODES = {"standard": 0, "easy": 1}


class NjodeError(RuntimeError):
    pass


class Lib:
    """one loaded build of the njode_b200 C ABI"""

    def __init__(self, path):
        if not os.path.exists(path):
            raise NjodeError(
                "njode_b200: native library %s not found -- build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` (there is no CPU fallback)" % path)
        self.path = path
        self.dll = C.CDLL(path)
        d = self.dll
        d.njode_last_error.restype = C.c_char_p
        d.njode_abi_version.restype = C.c_int
        d.njode_plan.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT), C.c_int, C.POINTER(PlanT)]
        d.njode_forward.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT), C.c_void_p, C.c_void_p,
                                    C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(SavedT),
                                    C.c_void_p, C.c_void_p]
        d.njode_backward.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT), C.c_void_p,
                                     C.POINTER(SavedT), C.c_void_p, C.c_void_p, C.c_void_p,
                                     C.c_void_p, C.c_void_p]
        for f in (d.njode_plan, d.njode_forward, d.njode_backward):
            f.restype = C.c_int
        if d.njode_abi_version() != 6:
            raise NjodeError("njode_b200: ABI version mismatch in %s" % path)
        # device index build (absent from the host simulation: the CPU-only tests use schedule.build_index_torch)
        self.has_index = hasattr(d, "njode_build_index")
        if self.has_index:
            d.njode_index_workspace_bytes.argtypes = [C.c_int32, C.c_int32]
            d.njode_index_workspace_bytes.restype = C.c_int64
            d.njode_build_index.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_void_p, C.c_int32, C.c_int32,
                                            C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
            d.njode_build_index.restype = C.c_int
        # tensor-core path (absent from the host simulation used by the CPU-only tests)
        self.has_wide = hasattr(d, "njode_wide_forward")
        if self.has_wide:
            d.njode_wide_supported.argtypes = [C.POINTER(ModelT)]
            d.njode_wide_supported.restype = C.c_int
            d.njode_wide_workspace_bytes.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT)]
            d.njode_wide_workspace_bytes.restype = C.c_int64
            d.njode_wide_saved_bytes.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT)]
            d.njode_wide_saved_bytes.restype = C.c_int64
            d.njode_wide_forward.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT), C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.POINTER(SavedT), C.c_void_p, C.c_void_p, C.c_void_p]
            d.njode_wide_forward.restype = C.c_int
            d.njode_wide_backward.argtypes = [C.POINTER(ModelT), C.POINTER(BatchT), C.c_void_p, C.c_void_p,
                                              C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
            d.njode_wide_backward.restype = C.c_int
            d.njode_wide_get_timing.argtypes = [C.POINTER(C.c_float)] * 3
            d.njode_wide_get_timing_bwd.argtypes = [C.POINTER(C.c_float)] * 2

    def check(self, rc, what):
        if rc != 0:
            raise NjodeError("%s failed (%d): %s" % (what, rc, self.dll.njode_last_error().decode()))


_cuda_lib = None


def cuda_lib():
    global _cuda_lib
    if _cuda_lib is None:
        _cuda_lib = Lib(CUDA_LIB_PATH)
    return _cuda_lib


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


class _Cast:
    """a host array that is converted to ``dtype`` while it is copied into the pinned staging block (one pass over the
    data instead of astype + copy)"""
    __slots__ = ("src", "dtype", "nbytes")

    def __init__(self, src, dtype):
        self.src, self.dtype = np.ascontiguousarray(src), np.dtype(dtype)
        self.nbytes = self.src.size * self.dtype.itemsize

    def copy_into(self, dst_u8):
        np.copyto(dst_u8.view(self.dtype), self.src.reshape(-1), casting="unsafe")


class PreparedBatch:
    """device-resident inputs of one forward call + the ctypes batch structs (fwd / bwd)."""
    __slots__ = ("sched", "B", "N", "dev", "keep", "fwd", "bwd_loss", "bwd_all", "n_units",
                 "n_loss_units", "h2d_bytes", "get_loss", "return_path", "runner")


class Runner:
    """stages one batch on `device`, launches forward / backward through `lib`."""

    def __init__(self, lib, device):
        self.lib = lib
        self.device = torch.device(device)
        self.is_cuda = self.device.type == "cuda"
        self.sms = torch.cuda.get_device_properties(self.device).multi_processor_count if self.is_cuda else 4
        self._ws = None
        self._pin = None
        self._pin_event = None

    # -- buffers ----------------------------------------------------------------------------
    def _workspace(self, nbytes):
        if self._ws is None or self._ws.numel() < nbytes:
            self._ws = torch.empty(int(nbytes * 1.25) + 1024, dtype=torch.uint8, device=self.device)
        return self._ws

    def _staging(self, nbytes):
        if self._pin_event is not None:
            self._pin_event.synchronize()       # previous async copy out of the pinned block is done
        if self._pin is None or self._pin.numel() < nbytes:
            cap = int(nbytes * 1.5) + 4096
            self._pin = torch.empty(cap, dtype=torch.uint8, pin_memory=self.is_cuda)
        # a fresh device block per batch: the autograd graph may keep several batches alive
        return self._pin, torch.empty(nbytes, dtype=torch.uint8, device=self.device)

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream) if self.is_cuda else None

    def _guard(self):
        """every C-ABI call plans and launches on the CURRENT CUDA device (cudaGetDevice): make it this runner's device
        for the duration of the call, so a model on cuda:1 works while cuda:0 is current"""
        return torch.cuda.device(self.device) if self.is_cuda else contextlib.nullcontext()

    # -- batch staging ----------------------------------------------------------------------
    def prepare(self, times, time_ptr, X, obs_idx, delta_t, T, start_X, n_obs_ot, M, until_T,
                return_path, segments, d, batch_size_norm=None, path_id_offset=0):
        """stages one batch of the collate contract: the host builds only the Euler schedule (float64
        loop conditions of the reference) and ships the raw arrays in ONE pinned host->device copy; the
        per-path CSR and the sorted work units are built on the device (schedule.build_index_torch)."""
        sched = _sched.build_schedule(times, delta_t, T, until_T, return_path)
        B = int(start_X.shape[0])
        n_idx = int(obs_idx.numel()) if torch.is_tensor(obs_idx) else len(obs_idx)
        if torch.is_tensor(time_ptr) and time_ptr.device.type != "cpu":
            # a batch collated on the device (stock_model.DeviceDataset.collate(..., on_device=True)): time_ptr / obs_idx /
            # n_obs_ot never visit the host; K comes from len(times) and N from the (already sliced) row arrays
            tp = time_ptr.detach().to(self.device, torch.int32).contiguous()
            K = int(tp.numel()) - 1
            N = n_idx
        else:
            tp = np.ascontiguousarray(np.asarray(time_ptr, dtype=np.int64).astype(np.int32))
            K = len(tp) - 1
            N = int(tp[-1]) if len(tp) else 0
            if n_idx != N:
                raise AssertionError("len(obs_idx) != time_ptr[-1]")

        def stage_f32(t, shape):
            if t is None:
                return None
            if torch.is_tensor(t):
                if t.device.type != "cpu":
                    return t.detach().to(self.device, torch.float32).contiguous().reshape(shape)
                a = t.detach().numpy()
            else:
                a = np.asarray(t)
            if a.dtype != np.float32:
                return _Cast(a, np.float32)
            return np.ascontiguousarray(a).reshape(shape)

        if torch.is_tensor(obs_idx) and obs_idx.device.type != "cpu":
            obs_arr = obs_idx.detach().to(self.device)
        else:
            obs_arr = obs_idx.detach().numpy() if torch.is_tensor(obs_idx) else np.asarray(obs_idx)
        # small batches: the index arrays are cheaper to build with NumPy on the host (a dozen tensor-op
        # launches cost more than sorting a few thousand rows); large ones are built on the device
        mode = os.environ.get("NJODE_INDEX", "auto")
        host_index = (not torch.is_tensor(obs_arr)) and (mode == "host" or (mode == "auto" and N + B < 16384))
        if not torch.is_tensor(obs_arr):
            if host_index:
                obs_arr = np.ascontiguousarray(obs_arr.astype(np.int32, copy=False))
            elif obs_arr.dtype != np.int32:
                obs_arr = _Cast(obs_arr, np.int32)          # int64 -> int32 on the way into the staging block
        budget = float(B) * sched.S / (16.0 * self.sms * 12)
        T1, T2 = max(8, int(0.75 * budget)), max(4, int(0.4 * budget))
        host_idx = {}
        if host_index:
            path_ptr, path_rows, row_jump = _sched.build_csr(tp, obs_arr, B)
            units, n_loss = _sched.build_units(sched, path_ptr, path_rows, row_jump, B, segments)
            host_idx = {"path_ptr": path_ptr, "path_rows": path_rows, "row_jump": row_jump, "unit_desc": units.reshape(-1)}
            lens = (units[:, 2] - units[:, 1])
            st = [0, 0, 0, 0, 0, 0]
            if segments:
                st[:4] = [int(np.count_nonzero(lens[:n_loss] >= T1)), int(np.count_nonzero(lens[:n_loss] >= T2)),
                          int(np.count_nonzero(lens[n_loss:] >= T1)), int(np.count_nonzero(lens[n_loss:] >= T2))]
        arrays = {"X": stage_f32(X, (N, d)), "M": stage_f32(M, (N, d)),
                  "start_X": stage_f32(start_X, (B, d)), "n_obs_ot": stage_f32(n_obs_ot, (B,)),
                  "obs_idx": None if host_index else obs_arr, "time_ptr": tp, **host_idx,
                  "step_dt": sched.step_dt, "step_t": sched.step_t, "jump_step": sched.jump_step,
                  "jump_tau": sched.jump_tau, "step_event": sched.step_event, "jump_event": sched.jump_event}
        # one pinned staging block, one host->device copy
        offs, total = {}, 0
        for k, a in arrays.items():
            if a is None or torch.is_tensor(a):
                continue
            offs[k] = total
            total += (a.nbytes + 15) & ~15
        pin, dev = self._staging(max(total, 16))
        pin_np = pin.numpy()
        for k, o in offs.items():
            a = arrays[k]
            if isinstance(a, _Cast):
                a.copy_into(pin_np[o:o + a.nbytes])
            else:
                pin_np[o:o + a.nbytes] = a.view(np.uint8).reshape(-1)
        dev.copy_(pin[:dev.numel()], non_blocking=True)
        if self.is_cuda:
            self._pin_event = torch.cuda.Event()
            self._pin_event.record(torch.cuda.current_stream(self.device))
        base = dev.data_ptr()
        keep = [dev]

        def view_i32(k, n):
            a = arrays[k]
            if torch.is_tensor(a):
                return a
            return dev[offs[k]:offs[k] + 4 * n].view(torch.int32)

        # tile-height classes of the segment kernels: units at least T1 (T2) Euler steps long are marched in
        # the lowest (middle) tiles; thresholds follow the per-warp share of the batch's total work B * S
        index = {}
        if not host_index and self.is_cuda:
            # njode_build_index: ~10 launches (histogram, scan, three stable radix sorts, unit kernels) behind one call
            if not self.lib.has_index:
                raise NjodeError("njode_b200: libnjode_b200.so lacks njode_build_index -- rebuild it")
            obs_t = view_i32("obs_idx", N)
            if obs_t.dtype != torch.int32:
                obs_t = obs_t.to(torch.int32)
            n_u = (N + B) if segments else B
            out = torch.empty(B + 1 + 2 * N + 6 * n_u + 8, dtype=torch.int32, device=self.device)
            path_ptr, path_rows, row_jump = out[:B + 1], out[B + 1:B + 1 + N], out[B + 1 + N:B + 1 + 2 * N]
            unit_desc, stats = out[B + 1 + 2 * N:B + 1 + 2 * N + 6 * n_u], out[B + 1 + 2 * N + 6 * n_u:]
            wsb = int(self.lib.dll.njode_index_workspace_bytes(N, B))
            iws = getattr(self, "_iws", None)
            if iws is None or iws.numel() < wsb:
                iws = self._iws = torch.empty(int(wsb * 1.25) + 1024, dtype=torch.uint8, device=self.device)
            with self._guard():
                rc = self.lib.dll.njode_build_index(
                    _ptr(obs_t), N, _ptr(view_i32("time_ptr", K + 1)), K, _ptr(view_i32("jump_step", K)), B, sched.S,
                    1 if segments else 0, T1, T2, _ptr(path_ptr), _ptr(path_rows), _ptr(row_jump), _ptr(unit_desc),
                    _ptr(stats), _ptr(iws), iws.numel(), self._stream())
            self.lib.check(rc, "njode_build_index")
            index = {"path_ptr": path_ptr, "path_rows": path_rows, "row_jump": row_jump, "unit_desc": unit_desc}
            keep.extend([out, obs_t])
            n_loss = N if segments else B
            st = stats[:6].tolist()                              # the one small device->host read of staging
        elif not host_index:
            # the host simulation used by the CPU-only tests: same arrays from a handful of tensor ops
            path_ptr, path_rows, row_jump, unit_desc, n_loss, stats = _sched.build_index_torch(
                view_i32("obs_idx", N), view_i32("time_ptr", K + 1), view_i32("jump_step", K), B, sched.S, segments, T1, T2)
            index = {"path_ptr": path_ptr, "path_rows": path_rows, "row_jump": row_jump, "unit_desc": unit_desc}
            keep.extend(index.values())
            st = [int(v) for v in stats.cpu()]
        if st[5]:
            raise IndexError("obs_idx out of range")
        if st[4]:
            # the contract has at most one row per (time, path) (NJODE/data_utils.py:302-306)
            raise ValueError("a path has two observation rows at the same observation time")
        seg_n1, seg_n2 = (C.c_int32 * 2)(st[0], st[2]), (C.c_int32 * 2)(st[1], st[3])
        n_units = (N + B) if segments else B

        def p(k):
            if k in index:
                return C.c_void_p(index[k].data_ptr())
            a = arrays[k]
            if a is None:
                return None
            if torch.is_tensor(a):
                keep.append(a)
                return C.c_void_p(a.data_ptr())
            return C.c_void_p(base + offs[k])

        def make(n_units_):
            return BatchT(B=B, N=N, K=sched.K, S=sched.S, E=sched.E, n_loss_units=int(n_loss),
                          seg_n1=seg_n1, seg_n2=seg_n2,
                          batch_size_norm=int(batch_size_norm or B), path_id_offset=int(path_id_offset),
                          n_units=int(n_units_), unit_kind=1 if segments else 0, X=p("X"), M=p("M"), start_X=p("start_X"),
                          n_obs_ot=p("n_obs_ot"), path_ptr=p("path_ptr"), path_rows=p("path_rows"),
                          row_jump=p("row_jump"), step_dt=p("step_dt"), step_t=p("step_t"),
                          jump_step=p("jump_step"), jump_tau=p("jump_tau"), step_event=p("step_event"),
                          jump_event=p("jump_event"), unit_desc=p("unit_desc"))

        pb = PreparedBatch()
        pb.sched, pb.B, pb.N, pb.dev, pb.keep = sched, B, N, self.device, keep
        pb.n_units, pb.n_loss_units = n_units, n_loss
        pb.fwd = make(n_units)
        pb.bwd_loss = make(n_loss)       # backward without a gradient into hT: tails contribute nothing
        pb.bwd_all = pb.fwd
        pb.h2d_bytes = total
        return pb

    # -- launches ---------------------------------------------------------------------------
    def plan(self, model_t, batch_t):
        pl = PlanT()
        dev = self.device.index if self.is_cuda and self.device.index is not None else 0
        with self._guard():
            self.lib.check(self.lib.dll.njode_plan(C.byref(model_t), C.byref(batch_t), dev, C.byref(pl)), "njode_plan")
        return pl

    def wide_supported(self, model_t):
        """True when the tcgen05 tensor-core path (njode_wide_*) can run this model"""
        return bool(self.lib.has_wide and self.is_cuda and self.lib.dll.njode_wide_supported(C.byref(model_t)))

    def _wide_workspace(self, nbytes):
        ws = getattr(self, "_wws", None)
        if ws is None or ws.numel() < nbytes:
            ws = self._wws = torch.empty(int(nbytes * 1.1) + 4096, dtype=torch.uint8, device=self.device)
        return ws

    def forward_wide(self, model_t, pb, params, H, dout, get_loss, need_grad, fp32_backward=False):
        """same contract as ``forward`` on the tensor-core kernels (bf16 operands, fp32 accumulate / state).
        ``saved`` is a ("wide", blob) pair for ``backward_wide``, or -- with ``fp32_backward`` -- the fp32
        kernels' (h_hist, h_before, y_after) triple."""
        f32 = dict(dtype=torch.float32, device=self.device)
        dll = self.lib.dll
        nbytes = dll.njode_wide_workspace_bytes(C.byref(model_t), C.byref(pb.fwd))
        if nbytes < 0:
            self.lib.check(int(nbytes), "njode_wide_workspace_bytes")
        ws = self._wide_workspace(nbytes)
        hT = torch.empty(pb.B, H, **f32)
        loss = torch.zeros((), **f32) if get_loss else None
        saved_t, saved, blob = SavedT(), None, None
        if need_grad and fp32_backward:
            saved = (torch.empty(max(pb.sched.S, 1) * pb.B * H, **f32), torch.empty(max(pb.N, 1) * H, **f32),
                     torch.empty(max(pb.N, 1) * dout, **f32))
            saved_t = SavedT(*[C.c_void_p(t.data_ptr()) for t in saved])
        elif need_grad:
            sbytes = dll.njode_wide_saved_bytes(C.byref(model_t), C.byref(pb.fwd))
            if sbytes < 0:
                self.lib.check(int(sbytes), "njode_wide_saved_bytes")
            blob = torch.empty(int(sbytes), dtype=torch.uint8, device=self.device)
            saved = ("wide", blob)
        with self._guard():
            rc = dll.njode_wide_forward(C.byref(model_t), C.byref(pb.fwd), _ptr(params), _ptr(hT), _ptr(loss),
                                        C.byref(saved_t), _ptr(blob), _ptr(ws), self._stream())
        self.lib.check(rc, "njode_wide_forward")
        return hT, loss, None, None, saved

    def backward_wide(self, model_t, pb, params, blob, grad_loss, grad_hT):
        dll = self.lib.dll
        nbytes = dll.njode_wide_workspace_bytes(C.byref(model_t), C.byref(pb.fwd))
        ws = self._wide_workspace(nbytes)
        grads = torch.empty_like(params)
        with self._guard():
            rc = dll.njode_wide_backward(C.byref(model_t), C.byref(pb.fwd), _ptr(params), _ptr(blob), _ptr(grad_loss),
                                         _ptr(grad_hT), _ptr(grads), _ptr(ws), self._stream())
        self.lib.check(rc, "njode_wide_backward")
        return grads

    def forward(self, model_t, pb, params, H, dout, get_loss, need_grad, recompute="auto"):
        """``recompute``: "on" / "off" / "auto" -- when the backward of this (model, batch) can recompute its forward
        from the checkpoints at the observation times (njode_plan: recompute_bytes > 0; the segment kernels), "on" saves
        NOTHING for it (no [S, B, H] history: memory independent of the batch size, ~15 % more backward time); "auto"
        does so when the history would exceed RECOMPUTE_AUTO_BYTES"""
        f32 = dict(dtype=torch.float32, device=self.device)
        pl = self.plan(model_t, pb.fwd)
        ws = self._workspace(pl.workspace_bytes)
        hT = torch.empty(pb.B, H, **f32)
        loss = torch.zeros((), **f32) if get_loss else None
        E = pb.sched.E
        path_h = torch.empty(E, pb.B, H, **f32) if E else None
        path_y = torch.empty(E, pb.B, dout, **f32) if E else None
        saved_t, saved = SavedT(), None
        hist_bytes = 4 * max(pb.sched.S, 1) * pb.B * H
        if need_grad and pl.recompute_bytes > 0 and (recompute == "on" or (recompute == "auto" and hist_bytes > RECOMPUTE_AUTO_BYTES)):
            saved = ()                       # the backward recomputes: nothing to keep
        elif need_grad:
            saved = (torch.empty(max(pb.sched.S, 1) * pb.B * H, **f32),
                     torch.empty(max(pb.N, 1) * H, **f32), torch.empty(max(pb.N, 1) * dout, **f32))
            if 0 < pl.act_bytes <= SAVE_ACTIVATIONS_MAX_BYTES and os.environ.get("NJODE_SAVE_ACTIVATIONS", "1") != "0":
                saved = saved + (torch.empty(pl.act_bytes // 4, **f32),)
            saved_t = SavedT(*[C.c_void_p(t.data_ptr()) for t in saved])
        with self._guard():
            rc = self.lib.dll.njode_forward(C.byref(model_t), C.byref(pb.fwd), _ptr(params), _ptr(hT),
                                            _ptr(loss), _ptr(path_h), _ptr(path_y), C.byref(saved_t),
                                            _ptr(ws), self._stream())
        self.lib.check(rc, "njode_forward")
        return hT, loss, path_h, path_y, saved

    def backward(self, model_t, pb, params, saved, grad_loss, grad_hT):
        # without a gradient into hT the tail units (after a path's last observation) contribute nothing
        bt = pb.bwd_all if grad_hT is not None else pb.bwd_loss
        pl = self.plan(model_t, bt)
        ws = self._workspace(pl.workspace_bytes)
        grads = torch.empty_like(params)
        saved_t = SavedT(*[C.c_void_p(t.data_ptr()) for t in saved]) if saved else SavedT()
        with self._guard():
            rc = self.lib.dll.njode_backward(C.byref(model_t), C.byref(bt), _ptr(params), C.byref(saved_t),
                                             _ptr(grad_loss), _ptr(grad_hT), _ptr(grads), _ptr(ws),
                                             self._stream())
        self.lib.check(rc, "njode_backward")
        return grads


_runners = {}


def cuda_runner(device):
    device = torch.device(device)
    if device.type != "cuda":
        raise NjodeError("njode_b200 runs on CUDA devices only; model parameters are on %s "
                         "(call model.to('cuda')) -- there is no CPU fallback" % device)
    if not torch.cuda.is_available():
        raise NjodeError("njode_b200: no CUDA device available")
    idx = device.index if device.index is not None else torch.cuda.current_device()
    r = _runners.get(idx)
    if r is None:
        r = _runners[idx] = Runner(cuda_lib(), torch.device("cuda", idx))
    return r
