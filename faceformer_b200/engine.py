"""Host-side owner of one libffb200 handle (one per device per process).

PyTorch is plumbing here: it supplies device memory (``torch.Tensor.data_ptr``)
and the current CUDA stream; all arithmetic happens inside libffb200.so.
Inputs may be CUDA tensors (device path: what ``Trainer.forward`` hands the model
after Lightning moved the batch, trainer.py:27-28) or host arrays / pinned tensors
(host path: the library does the H2D/D2H copies itself).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import lib as _lib
from .config import MODE_PARALLEL, MODE_SEQ2SEQ, ModelConfig
from .lib import FFB_DEVICE, FFB_HOST, FFBError
from .synth import state_dict_names


def pack_state_dict(sd, cfg: ModelConfig, mode: int) -> np.ndarray:
    """Strictly validate a reference ``state_dict`` (names, shapes) and
    concatenate its float tensors in state_dict order into the fp32 blob ``ffb_load_weights``
    expects.  Accepts numpy arrays or torch tensors; a leading ``model.`` prefix (Lightning
    checkpoints, trainer.py:20) is stripped."""
    clean = {}
    for k, v in sd.items():
        k = k[6:] if k.startswith("model.") else k
        clean[k] = v
    expected = state_dict_names(cfg, mode)
    names = {n for n, _, _ in expected}
    missing = [n for n in names if n not in clean]
    unexpected = [k for k in clean if k not in names]
    if missing or unexpected:
        raise FFBError(f"state_dict mismatch: missing {sorted(missing)[:5]} unexpected {sorted(unexpected)[:5]}")
    parts = []
    for name, shape, dt in expected:
        v = clean[name]
        if hasattr(v, "detach"):
            v = v.detach().cpu().numpy()
        v = np.asarray(v)
        if tuple(v.shape) != tuple(shape):
            raise FFBError(f"state_dict[{name!r}] has shape {tuple(v.shape)}, expected {tuple(shape)}")
        if dt == "i8":
            if not np.array_equal(v.reshape(-1), np.arange(shape[1])):
                raise FFBError(f"{name} must be arange (embedding.py:98-99)")
            continue
        parts.append(np.ascontiguousarray(v, dtype=np.float32).reshape(-1))
    return np.concatenate(parts)


def _ptr(a):
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _is_cuda(a) -> bool:
    return hasattr(a, "is_cuda") and bool(a.is_cuda)


class Engine:
    """``model(batch)`` for one device.  Not thread-safe (one handle, stream-ordered calls)."""

    def __init__(self, cfg: ModelConfig, mode: int, device: int = 0):
        self.cfg, self.mode, self.device = cfg, mode, int(device)
        self._lib = _lib.load()
        c = _lib.ffb_config(abi_version=_lib.FFB_ABI_VERSION, mode=mode, num_model=cfg.num_model,
                            num_head=cfg.num_head, num_feedforward=cfg.num_feedforward,
                            num_encoder_layers=cfg.num_encoder_layers, num_decoder_layers=cfg.num_decoder_layers,
                            in_dim=cfg.in_dim, num_lines=cfg.num_lines, num_token=cfg.num_token,
                            seq_len=cfg.seq_len(mode), device=self.device)
        self._c = c
        self._h = C.c_void_p()
        st = self._lib.ffb_create(C.byref(c), C.byref(self._h))
        if st != 0:
            raise FFBError(f"ffb_create failed ({st}): {_lib.last_error(None)}")
        self.weights_loaded = False

    # -- lifecycle ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.ffb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        _lib.check(st, self._h)

    def _stream(self):
        import torch
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    # -- weights --------------------------------------------------------------------------------
    def weight_count(self) -> int:
        return int(self._lib.ffb_weight_count(C.byref(self._c)))

    def load_state_dict(self, sd):
        blob = pack_state_dict(sd, self.cfg, self.mode)
        self.load_blob(blob)

    def load_blob(self, blob):
        """blob: fp32 numpy array (host) or CUDA tensor in ffb_load_weights order."""
        loc = FFB_DEVICE if _is_cuda(blob) else FFB_HOST
        n = blob.numel() if hasattr(blob, "numel") else blob.size
        if loc == FFB_HOST:
            blob = np.ascontiguousarray(blob, dtype=np.float32)
        self._check(self._lib.ffb_load_weights(self._h, _ptr(blob), n, loc, self._stream()))
        self.weights_loaded = True

    def set_option(self, opt: int, value: int):
        self._check(self._lib.ffb_set_option(self._h, opt, int(value)))

    # -- the path -------------------------------------------------------------------------------
    def _prep_inputs(self, coords, pad_mask, num_input):
        import torch
        dev = _is_cuda(coords)
        if dev:
            coords = coords.contiguous().float()
            pad_mask = pad_mask.to(device=coords.device).contiguous().to(torch.uint8)
            if num_input is not None:
                num_input = torch.as_tensor(num_input).to(device=coords.device, dtype=torch.int64).contiguous()
            n = coords.shape[0]
        else:
            if hasattr(coords, "numpy"):
                coords = coords.numpy()
            if hasattr(pad_mask, "numpy"):
                pad_mask = pad_mask.numpy()
            if num_input is not None and hasattr(num_input, "numpy"):
                num_input = num_input.numpy()
            coords = np.ascontiguousarray(coords, dtype=np.float32)
            pad_mask = np.ascontiguousarray(pad_mask).astype(np.uint8, copy=False)
            if num_input is not None:
                num_input = np.ascontiguousarray(num_input, dtype=np.int64)
            n = coords.shape[0]
        exp = (n, self.cfg.num_lines)
        if tuple(coords.shape[:2]) != exp or int(np.prod(coords.shape[2:])) != self.cfg.in_dim:
            raise FFBError(f"input has shape {tuple(coords.shape)}, expected [N,{self.cfg.num_lines},P,D] with P*D={self.cfg.in_dim}")
        if tuple(pad_mask.shape) != exp:
            raise FFBError(f"input_mask has shape {tuple(pad_mask.shape)}, expected {exp}")
        if self.mode == MODE_PARALLEL and num_input is None:
            raise FFBError("num_input is required for SurfaceFormer_Parallel (model_para.py:187)")
        if self.mode == MODE_SEQ2SEQ:
            num_input = None
        return coords, pad_mask, num_input, n, (FFB_DEVICE if dev else FFB_HOST)

    def encode(self, coords, pad_mask, num_input=None):
        coords, pad_mask, num_input, n, loc = self._prep_inputs(coords, pad_mask, num_input)
        self._keep = (coords, pad_mask, num_input)          # keep alive until the stream consumed them
        self._check(self._lib.ffb_encode(self._h, _ptr(coords), _ptr(pad_mask), _ptr(num_input), n, loc, self._stream()))
        self._loc, self._n = loc, n
        return self.batch_info()

    def batch_info(self) -> dict:
        n, f = C.c_int32(), C.c_int32()
        b, be, r = C.c_int64(), C.c_int64(), C.c_int64()
        self._check(self._lib.ffb_batch_info(self._h, C.byref(n), C.byref(f), C.byref(b), C.byref(be), C.byref(r)))
        return dict(N=n.value, F=f.value, B=b.value, B_eff=be.value, R=r.value)

    def _alloc(self, shape, dtype, loc):
        import torch
        if loc == FFB_DEVICE:
            return torch.empty(shape, dtype=dtype, device=f"cuda:{self.device}")
        return np.empty(shape, dtype={torch.int64: np.int64, torch.float32: np.float32}[dtype])

    def decode_greedy(self, want_steps: bool = True, out=None):
        """Returns (predict, steps).  predict is int64 [N,F,T] / [N,T] on the inputs' side (CUDA tensor or numpy)."""
        import torch
        info = self.batch_info()
        T = self.cfg.seq_len(self.mode)
        shape = (info["N"], info["F"], T) if self.mode == MODE_PARALLEL else (info["N"], T)
        predict = out if out is not None else self._alloc(shape, torch.int64, self._loc)
        steps = C.c_int32(-1)
        self._check(self._lib.ffb_decode_greedy(self._h, _ptr(predict), self._loc,
                                                C.byref(steps) if want_steps else None, self._stream()))
        return predict, (steps.value if want_steps else None)

    def forward_eval(self, coords, pad_mask, num_input=None, want_steps: bool = True, out=None):
        self.encode(coords, pad_mask, num_input)
        return self.decode_greedy(want_steps, out)

    def forward_train(self, coords, pad_mask, num_input, label, label_mask, want_embedding: bool = False):
        """Teacher-forced forward pass (forward only).  want_embedding: also return the reference's `embedding` [N, L, E] (before its
        replication per anchor slot) with the rows of padded edges computed as the reference does -- compute_loss's softmax runs over them.  Parallel model (model_para.py:99-171, scheduled_sampling_ratio = 0): label int64 /
        label_mask bool [N, rows >= F, T] -> pointer f32 [N * F, T - 1, E], F = max(num_input).  Seq2seq model (model.py:98-157): label /
        label_mask [N, T] -> pointer [N, T - 1, E].  The result lives on the inputs' side (CUDA tensor or numpy)."""
        import torch
        coords, pad_mask, num_input, n, loc = self._prep_inputs(coords, pad_mask, num_input)
        if loc == FFB_DEVICE:
            label = label.contiguous().to(torch.int64)
            label_mask = label_mask.contiguous().to(torch.uint8)
        else:
            label = np.ascontiguousarray(label, dtype=np.int64)
            label_mask = np.ascontiguousarray(label_mask, dtype=np.uint8)
        T = self.cfg.seq_len(self.mode)
        if self.mode == MODE_PARALLEL:
            f = int(num_input.max().item()) if loc == FFB_DEVICE else int(np.max(num_input))
            want = 3
        else:
            f, want = 1, 2
        if label.ndim != want or label.shape[0] != n or label.shape[-1] != T or tuple(label_mask.shape) != tuple(label.shape):
            raise FFBError(f"label / label_mask must be [N={n}, {'rows, ' if want == 3 else ''}T={T}], got {tuple(label.shape)} / {tuple(label_mask.shape)}")
        rows = int(label.shape[1]) if want == 3 else 1
        out = self._alloc((n * f, T - 1, self.cfg.num_model), torch.float32, loc)
        emb = self._alloc((n, self.cfg.mem_len, self.cfg.num_model), torch.float32, loc) if want_embedding else None
        self._keep = (coords, pad_mask, num_input, label, label_mask)
        self._check(self._lib.ffb_forward_train(self._h, _ptr(coords), _ptr(pad_mask), _ptr(num_input), n, _ptr(label), _ptr(label_mask),
                                                rows, _ptr(out), _ptr(emb) if want_embedding else None, loc, self._stream()))
        self._loc, self._n = loc, n
        return (out, emb) if want_embedding else out

    # -- the step before the path ------------------------------------------------
    def featurize(self, wireframes, device: bool = True):
        """wireframes: list (one per wireframe) of lists of edges, an edge = sequence of (x, y) points -- the `edges` entry of the
        dataset JSON (datasets/data_para.py:59-68).  Returns (input [N, num_lines, P, 2] f32, input_mask [N, num_lines] bool,
        num_input [N] i64) as torch tensors on the engine's GPU (device=True) or numpy arrays computed through host buffers."""
        import torch
        pts, eoff, woff = self._ragged(wireframes)
        n, nl, P = len(woff) - 1, self.cfg.num_lines, self.cfg.num_points_per_line
        if device:
            dev = torch.device("cuda", self.device)
            t_pts, t_e, t_w = (torch.from_numpy(x).to(dev) for x in (pts, eoff, woff))
            out = torch.empty((n, nl, P, 2), dtype=torch.float32, device=dev)
            mask = torch.empty((n, nl), dtype=torch.uint8, device=dev)
            ni = torch.empty((n,), dtype=torch.int64, device=dev)
            self._check(self._lib.ffb_featurize(self._h, _ptr(t_pts), _ptr(t_e), _ptr(t_w), n, _ptr(out), _ptr(mask), _ptr(ni),
                                                FFB_DEVICE, self._stream()))
            return out, mask.bool(), ni
        out = np.empty((n, nl, P, 2), np.float32)
        mask = np.empty((n, nl), np.uint8)
        ni = np.empty((n,), np.int64)
        self._check(self._lib.ffb_featurize(self._h, _ptr(pts), _ptr(eoff), _ptr(woff), n, _ptr(out), _ptr(mask), _ptr(ni),
                                            FFB_HOST, self._stream()))
        return out, mask.astype(bool), ni

    # -- the step after the path --------------------------------------------------
    @staticmethod
    def flatten_wireframes(wireframes):
        """Nested `edges` lists -> the flat (points [n, 2] f64, edge offsets, wireframe offsets) arrays the C ABI takes.  This Python loop
        costs as much as the reference's own per-edge numpy featurisation (profiles/probe_f1_f2_r2.json): do it ONCE per dataset (the JSON
        is static) and hand the tuple to featurize / parse_faces instead of the nested lists."""
        return Engine._ragged(wireframes)

    @staticmethod
    def _ragged(wireframes):
        if isinstance(wireframes, tuple) and len(wireframes) == 3 and all(isinstance(a, np.ndarray) for a in wireframes):
            return wireframes                                  # already flat (flatten_wireframes)
        pts, eoff, woff = [], [0], [0]
        for edges in wireframes:
            for e in edges:
                a = np.asarray(e, dtype=np.float64).reshape(-1, 2)
                pts.append(a)
                eoff.append(eoff[-1] + a.shape[0])
            woff.append(woff[-1] + len(edges))
        pts = np.ascontiguousarray(np.concatenate(pts, 0) if pts else np.zeros((0, 2), np.float64))
        return pts, np.asarray(eoff, np.int64), np.asarray(woff, np.int64)

    def parse_faces(self, predict, wireframes, tol: float = 2e-4, check_enclosed: bool = True):
        """predict: int64 [N, F, T] (CUDA tensor or numpy) as returned by forward_eval; wireframes: the raw `edges` lists.
        Returns one list per wireframe of (face_type, loops) -- loops = tuple of tuples of edge indices, rolled and ordered as
        filter_faces_by_encloseness does (post_processing.py:8-20); with check_enclosed=False (face_type, indices) as
        Trainer.parse_parallel_faces returns them (trainer.py:196-206)."""
        import torch
        pts, eoff, woff = self._ragged(wireframes)
        dev = _is_cuda(predict)
        n, f, t = predict.shape
        if t != self.cfg.seq_len(self.mode):
            raise FFBError(f"predict has T={t}, the engine was created for T={self.cfg.seq_len(self.mode)}")
        if dev:
            d = predict.device
            predict = predict.contiguous()
            ins = [torch.from_numpy(x).to(d) for x in (pts, eoff, woff)]
            valid = torch.empty((n, f), dtype=torch.uint8, device=d)
            ft, nl, ni = (torch.empty((n, f), dtype=torch.int32, device=d) for _ in range(3))
            ll, idx = (torch.empty((n, f, t), dtype=torch.int32, device=d) for _ in range(2))
            loc = FFB_DEVICE
        else:
            predict = np.ascontiguousarray(predict, dtype=np.int64)
            ins = [pts, eoff, woff]
            valid = np.empty((n, f), np.uint8)
            ft, nl, ni = (np.empty((n, f), np.int32) for _ in range(3))
            ll, idx = (np.empty((n, f, t), np.int32) for _ in range(2))
            loc = FFB_HOST
        self._check(self._lib.ffb_parse_faces(self._h, _ptr(predict), n, f, _ptr(ins[0]), _ptr(ins[1]), _ptr(ins[2]), float(tol),
                                              1 if check_enclosed else 0, _ptr(valid), _ptr(ft), _ptr(nl), _ptr(ll), _ptr(idx), _ptr(ni),
                                              loc, self._stream()))
        if dev:
            valid, ft, nl, ni, ll, idx = (x.cpu().numpy() for x in (valid, ft, nl, ni, ll, idx))
        out = []
        for w in range(n):
            faces = []
            for s in np.nonzero(valid[w])[0]:
                ind = idx[w, s, :ni[w, s]].tolist()
                if check_enclosed:
                    loops, pos = [], 0
                    for le in ll[w, s, :nl[w, s]].tolist():
                        loops.append(tuple(ind[pos:pos + le]))
                        pos += le
                    faces.append((int(ft[w, s]), tuple(loops)))
                else:
                    faces.append((int(ft[w, s]), tuple(ind)))
            out.append(faces)
        return out

    # -- parity hooks ---------------------------------------------------------------------------
    def get_memory(self):
        import torch
        info = self.batch_info()
        out = self._alloc((info["N"], self.cfg.mem_len, self.cfg.num_model), torch.float32, self._loc)
        self._check(self._lib.ffb_get_memory(self._h, _ptr(out), self._loc, self._stream()))
        return out

    def set_memory(self, memory):
        """Error-budget hook: replace the encoder memory [N, L, E] (CUDA tensor or numpy) and recompute the cross K / V cache."""
        dev = _is_cuda(memory)
        memory = memory.contiguous().float() if dev else np.ascontiguousarray(memory, dtype=np.float32)
        self._check(self._lib.ffb_set_memory(self._h, _ptr(memory), FFB_DEVICE if dev else FFB_HOST, self._stream()))

    def get_last_logits(self):
        import torch
        info = self.batch_info()
        out = self._alloc((info["B"], self.cfg.mem_len), torch.float32, self._loc)
        self._check(self._lib.ffb_get_last_logits(self._h, _ptr(out), self._loc, self._stream()))
        return out

    def get_last_pointer(self):
        import torch
        info = self.batch_info()
        T = self.cfg.seq_len(self.mode)
        buf = self._alloc((info["B_eff"], T - 1, self.cfg.num_model), torch.float32, self._loc)
        P = C.c_int32()
        self._check(self._lib.ffb_get_last_pointer(self._h, _ptr(buf), C.byref(P), self._loc, self._stream()))
        flat = buf.reshape(-1)[: info["B_eff"] * P.value * self.cfg.num_model]
        return flat.reshape(info["B_eff"], P.value, self.cfg.num_model)

    def forced_prefix_logits(self, prefix):
        """prefix int64 [P,B] (numpy or CUDA tensor) -> logits [B,L]."""
        import torch
        info = self.batch_info()
        dev = _is_cuda(prefix)
        if dev:
            prefix = prefix.contiguous().to(torch.int64)
        else:
            prefix = np.ascontiguousarray(prefix, dtype=np.int64)
        if tuple(prefix.shape[1:]) != (info["B"],):
            raise FFBError(f"prefix has shape {tuple(prefix.shape)}, expected [P,{info['B']}]")
        loc = FFB_DEVICE if dev else FFB_HOST
        out = self._alloc((info["B"], self.cfg.mem_len), torch.float32, loc)
        self._check(self._lib.ffb_forced_prefix_logits(self._h, _ptr(prefix), int(prefix.shape[0]), _ptr(out), loc, self._stream()))
        return out

    def fp16_fallbacks(self) -> int:
        return int(self._lib.ffb_fp16_fallbacks(self._h))

    def get_beams(self, width: int):
        """After a decode with FFB_OPT_BEAM = width: (beams int64 [N,F,W,T], scores float64 [N,F,W]), best hypothesis first."""
        import torch
        info = self.batch_info()
        T = self.cfg.seq_len(self.mode)
        if self._loc == FFB_DEVICE:
            beams = torch.empty((info["N"], info["F"], width, T), dtype=torch.int64, device=f"cuda:{self.device}")
            scores = torch.empty((info["N"], info["F"], width), dtype=torch.float64, device=f"cuda:{self.device}")
        else:
            beams = np.empty((info["N"], info["F"], width, T), np.int64)
            scores = np.empty((info["N"], info["F"], width), np.float64)
        self._check(self._lib.ffb_get_beams(self._h, _ptr(beams), _ptr(scores), self._loc, self._stream()))
        return beams, scores

    # -- one batch split over several GPUs ------------------------------------------------------
    def stop_exchange_export(self) -> bytes:
        buf = C.create_string_buffer(64)
        self._check(self._lib.ffb_stop_exchange_export(self._h, buf))
        return buf.raw

    def stop_exchange_connect(self, rank: int, world: int, handles):
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        self._check(self._lib.ffb_stop_exchange_connect(self._h, rank, world, C.c_char_p(blob)))

    def stop_exchange_disconnect(self):
        self._check(self._lib.ffb_stop_exchange_disconnect(self._h))

    def overflowed(self) -> bool:
        """After an asynchronous decode (want_steps=False on CUDA tensors): did an activation leave the fp16 range?  True means the
        predictions are invalid, the handle now runs in bf16x3, and the batch must be run again."""
        flag = C.c_int32(0)
        self._check(self._lib.ffb_overflowed(self._h, C.byref(flag), self._stream()))
        return bool(flag.value)

    def steps_launched(self) -> int:
        return int(self._lib.ffb_steps_launched(self._h))

    def used_persistent(self) -> bool:
        """Did the last decode run inside the persistent cooperative kernel (FFB_OPT_PERSISTENT)?"""
        return bool(self._lib.ffb_used_persistent(self._h))

    def kernel_launches(self) -> int:
        return int(self._lib.ffb_kernel_launches(self._h))

    def phase_times(self):
        arr = (C.c_float * 2)()
        self._check(self._lib.ffb_phase_times(self._h, arr, 2))
        return float(arr[0]), float(arr[1])

    def profile_read(self) -> dict:
        """class -> dict(ms, flops, launches) since profiling was enabled / last read."""
        n = len(_lib.PROFILE_CLASSES)
        ms, fl, ln = (C.c_float * n)(), (C.c_double * n)(), (C.c_int64 * n)()
        self._check(self._lib.ffb_profile_read(self._h, n, ms, fl, ln))
        return {c: dict(ms=float(ms[i]), flops=float(fl[i]), launches=int(ln[i])) for i, c in enumerate(_lib.PROFILE_CLASSES)}

    # -- op-level hooks (CUDA tensors)------------------------------------------------------------
    def op_linear(self, A, W, bias=None, R=None, pos=None, pos_mod=0, pos_cols=0, relu=False):
        import torch
        M, K = A.shape
        N = W.shape[0]
        Cc = torch.empty((M, N), dtype=torch.float32, device=A.device)
        self._check(self._lib.ffb_op_linear(self._h, _ptr(A), _ptr(W), _ptr(bias), _ptr(R), _ptr(pos), pos_mod, pos_cols,
                                            _ptr(Cc), M, N, K, int(relu), self._stream()))
        return Cc

    def op_linear_tc(self, A, W, bias=None, R=None, relu=False, via_split=False):
        import torch
        M, K = A.shape
        N = W.shape[0]
        Cc = torch.empty((M, N), dtype=torch.float32, device=A.device)
        self._check(self._lib.ffb_op_linear_tc(self._h, _ptr(A), _ptr(W), _ptr(bias), _ptr(R), _ptr(Cc), M, N, K, int(relu),
                                               int(via_split), self._stream()))
        return Cc

    def op_layernorm(self, x, gamma, beta):
        import torch
        y = torch.empty_like(x)
        self._check(self._lib.ffb_op_layernorm(self._h, _ptr(x), _ptr(gamma), _ptr(beta), _ptr(y), x.shape[0], x.shape[1], self._stream()))
        return y

    def op_attention(self, kind, q, k, v, G, nq, nk):
        import torch
        H = self.cfg.num_head
        out = torch.empty((G * nq, H * 64), dtype=torch.float32, device=q.device)
        self._check(self._lib.ffb_op_attention(self._h, kind, _ptr(q), q.stride(0), _ptr(k), _ptr(v), k.stride(0), _ptr(out),
                                               G, nq, nk, H, self._stream()))
        return out
