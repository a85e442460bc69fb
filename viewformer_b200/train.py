"""Codebook training step — forward + backward + Adam + data-parallel gradient exchange, on libvf_b200 kernels.

Reference: viewformer/models/vqgan_th.py:395-423 (forward, _compute_loss, training_step), :443-445 (Adam, betas (0.5, 0.9)),
models/utils_th.py:32-68 (QuantizeEMA: straight-through estimator, codebook moved by EMA — not by the gradient — with the
statistics all-reduced across ranks), train/train_codebook_th.py:39-41 (DDP: gradients averaged over ranks).  fp32, as the
reference requires (vqgan_th.py:326); LPIPS is not available offline, so ``perceptual_weight`` must be 0.

    trainer = VQGANTrainer(VQGAN(cfg, precision="fp32").load_state_dict(sd))
    loss = trainer.training_step(x)        # x f32 NCHW in [-1, 1]: forward, backward, gradient all-reduce, Adam, codebook EMA
    trainer.export_state_dict()            # reference-keyed weights after the step

Data layout: every trainable tensor lives in ONE flat fp32 buffer (kernel layouts: conv [kh*kw*Cin, Cout], dense [out, in]) with a
twin flat gradient buffer ordered by backward completion (decoder.conv_out first, encoder.conv_in last), so that
  * the data-parallel exchange is a handful of large NCCL all-reduces over contiguous buckets, each launched (async) the moment
    the backward pass has produced its last gradient — the transfers ride under the remaining backward kernels;
  * Adam is one kernel launch over the whole model.
Activations needed by the backward pass are kept on a tape; GroupNorm+swish outputs are recomputed from the saved statistics.

Arithmetic: fp32 throughout, as the reference requires.  The 3x3 stride-1 convolutions with tensor-core-sized channel counts run
their forward pass and their data gradient on the exact split-fp16 tcgen05 kernels (fp32-faithful results from three fp16 MMA passes,
DESIGN.md 5.3; ``VF_TRAIN_TC=0`` keeps everything on the CUDA cores), and their weight gradient as nine exact GEMMs over the pixel axis
(``_lib.conv_wgrad_tc``) when both channel counts are multiples of 128; strided / upsampling convs, 1x1 layers and the remaining weight
gradients use the fp32 CUDA-core kernels.
"""
import math
import os

import torch

from . import _lib as L
from .ops import linear


class _P:
    """One trainable tensor: kernel-layout view into the flat parameter buffer + its gradient view, and how to export it."""

    def __init__(self, name, tensor, setter, kind, part=None, cin=None):
        self.name, self.tensor, self.setter, self.kind, self.part, self.cin = name, tensor, setter, kind, part, cin
        self.grad = None


class VQGANTrainer:
    def __init__(self, model, lr=None, betas=(0.5, 0.9), eps=1e-8, bucket_bytes=64 << 20, process_group=None):
        if model.enc_prec.name != "fp32" or model.dec_prec.name != "fp32":
            raise ValueError("VQGANTrainer runs the fp32 path (the reference asserts no mixed precision, vqgan_th.py:326): build the model "
                             "with precision='fp32'")
        if model.config.perceptual_weight != 0:
            raise NotImplementedError("LPIPS (VGG16 weights) is not available offline: set perceptual_weight=0")
        model._need_weights()
        self.model, self.cfg = model, model.config
        self.lr = float(model.learning_rate if lr is None else lr)
        self.betas, self.eps, self.step_count = betas, eps, 0
        self.group = process_group
        self.bucket_bytes = bucket_bytes
        self._collect_params()
        self._flatten()
        self.last = {}
        self.use_tc = os.environ.get("VF_TRAIN_TC", "1") != "0"
        self._wsplit = {}

    # ------------------------------------------------------------------ parameter registry
    def _collect_params(self):
        w = self.model._w
        ps = []

        def conv(name, cw):
            key = "w_kn" if hasattr(cw, "w_kn") else None
            if key is None:
                raise RuntimeError(f"{name}: tensor-core weight layout in an fp32 model")
            ps.append(_P(name + ".weight", cw.w_kn, lambda t, cw=cw: setattr(cw, "w_kn", t), "conv", cin=cw.cin))
            ps.append(_P(name + ".bias", cw.bias, lambda t, cw=cw: setattr(cw, "bias", t), "vec"))

        def lin(name, ln, parts=None):
            ps.append(_P(name + ".weight", ln.w, lambda t, ln=ln: setattr(ln, "w", t), "dense", parts))
            ps.append(_P(name + ".bias", ln.b, lambda t, ln=ln: setattr(ln, "b", t), "vec", parts))

        def norm(name, d, key):
            ps.append(_P(name + ".weight", d[key][0], lambda t, d=d, key=key: d.__setitem__(key, (t, d[key][1])), "vec"))
            ps.append(_P(name + ".bias", d[key][1], lambda t, d=d, key=key: d.__setitem__(key, (d[key][0], t)), "vec"))

        def rb(name, r):
            norm(name + ".norm1", r, "n1"); conv(name + ".conv1", r["c1"]); norm(name + ".norm2", r, "n2"); conv(name + ".conv2", r["c2"])
            if "sc" in r:
                lin(name + ".nin_shortcut", r["sc"])

        def at(name, a):
            norm(name + ".norm", a, "norm")
            lin(name + ".qk", a["qk"], parts=(name + ".q", name + ".k"))
            lin(name + ".v", a["v"]); lin(name + ".proj_out", a["proj"])

        e, d = w["enc"], w["dec"]
        conv("encoder.conv_in", e["conv_in"])
        for lv, lvw in enumerate(e["levels"]):
            for b, r in enumerate(lvw["blocks"]):
                rb(f"encoder.down.{lv}.block.{b}", r)
                if lvw["attns"]:
                    at(f"encoder.down.{lv}.attn.{b}", lvw["attns"][b])
            if lvw["down"] is not None:
                conv(f"encoder.down.{lv}.downsample.conv", lvw["down"])
        rb("encoder.mid.block_1", e["mid1"]); at("encoder.mid.attn_1", e["mida"]); rb("encoder.mid.block_2", e["mid2"])
        norm("encoder.norm_out", e, "norm_out"); conv("encoder.conv_out", e["conv_out"])
        lin("quant_conv", w["quant_conv"])
        if self.model.quantizer == "commit":          # Quantize (utils_th.py:75-124): the codebook is an ordinary parameter
            q = w["q"]
            ps.append(_P("quantize.embeddings", q["emb"], lambda t, q=q: q.__setitem__("emb", t), "vec"))
        lin("post_quant_conv", w["post_quant_conv"])
        conv("decoder.conv_in", d["conv_in"])
        rb("decoder.mid.block_1", d["mid1"]); at("decoder.mid.attn_1", d["mida"]); rb("decoder.mid.block_2", d["mid2"])
        for lv in reversed(range(len(self.cfg.ch_mult))):
            lvw = d["levels"][lv]
            for b, r in enumerate(lvw["blocks"]):
                rb(f"decoder.up.{lv}.block.{b}", r)
                if lvw["attns"]:
                    at(f"decoder.up.{lv}.attn.{b}", lvw["attns"][b])
            if lvw["up"] is not None:
                conv(f"decoder.up.{lv}.upsample.conv", lvw["up"])
        norm("decoder.norm_out", d, "norm_out"); conv("decoder.conv_out", d["conv_out"])
        self.params = ps

    def _flatten(self):
        """Re-home every parameter in one flat buffer, in BACKWARD order; twin flat buffers for the gradient and Adam's moments."""
        dev = self.model.device
        order = list(reversed(self.params))
        offs, n = [], 0
        for p in order:
            offs.append(n)
            n += (p.tensor.numel() + 3) // 4 * 4            # 16-byte aligned views
        self.flat_p = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros_like(self.flat_p)
        self.flat_m = torch.zeros_like(self.flat_p)
        self.flat_v = torch.zeros_like(self.flat_p)
        for p, o in zip(order, offs):
            view = self.flat_p[o:o + p.tensor.numel()].view(p.tensor.shape)
            view.copy_(p.tensor)
            p.setter(view)
            p.tensor, p.offset = view, o
            p.grad = self.flat_g[o:o + p.tensor.numel()].view(p.tensor.shape)
        self.order = order
        # buckets: contiguous ranges of the flat gradient, closed after the parameter that pushes them past bucket_bytes
        self.buckets, start = [], 0
        for i, p in enumerate(order):
            end = p.offset + (p.tensor.numel() + 3) // 4 * 4
            if (end - start) * 4 >= self.bucket_bytes or i == len(order) - 1:
                self.buckets.append((start, end, p.name))
                start = end
        self._bucket_of = {}
        bi = 0
        for p in order:
            while p.offset >= self.buckets[bi][1]:
                bi += 1
            self._bucket_of[p.name] = bi
        self._bucket_size = [sum(1 for b in self._bucket_of.values() if b == i) for i in range(len(self.buckets))]
        self.model._refresh_decode_table()

    # ------------------------------------------------------------------ data-parallel exchange
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def _grad_ready(self, p):
        """Called when the backward pass has finished the gradient of ``p``: if that closes a bucket, start its all-reduce."""
        b = self._bucket_of[p.name]
        self._bucket_left[b] -= 1
        if self._bucket_left[b] == 0:
            self.launched.append(b)
            if self._world() > 1:
                import torch.distributed as dist
                s, e, _ = self.buckets[b]
                self._handles.append(dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        elif self._bucket_left[b] < 0:
            raise RuntimeError(f"gradient of {p.name} signalled twice")

    # ------------------------------------------------------------------ primitive forward / backward pairs
    def _pmap(self):
        return {p.name: p for p in self.params}

    def _gn_fw(self, x, nw):
        st = L.gn_mean_rstd(x)
        return st

    def _gn_apply(self, x, st, nw, swish):
        n, h, w, c = x.shape
        y = torch.empty_like(x)
        lib = L.load(True)
        L._check(lib.vf_groupnorm_apply(L._p(x), L.F32, L._p(st), L._p(nw[0]), L._p(nw[1]), n, h, w, c, 32, L.C.c_float(1e-6), 1, int(swish), 0,
                                        L._p(y), L.F32, L._stream()))
        return y

    # 3x3 stride-1 convolutions whose channel counts fit the tcgen05 tiles run on the EXACT split-fp16 tensor-core path (three fp16 MMA
    # passes, chunked accumulation: fp32-faithful results) in the forward pass and in the data gradient; everything else
    # (conv_in / conv_out, stride-2 and upsampling convs, 1x1 layers) and every weight gradient stays on the fp32 CUDA-core kernels.
    def _tc_ok(self, cw, stride, upsample):
        """3x3 stride-1 convs, incl. the Upsample convs (nearest x2, then a stride-1 conv on the doubled map: vqgan_th.py:29-32)."""
        return self.use_tc and cw.k == 3 and stride == 1 and cw.cin % 64 == 0 and cw.cout % 64 == 0

    def _split_weight(self, key, w_kn, n_out):
        """[K, n_out] fp32 (K = tap * C + c) -> split-fp16 [n_out, tap * 2C] for L.tc_conv; cached until the next optimizer step."""
        hit = self._wsplit.get(key)
        if hit is None:
            k = w_kn.shape[0]
            c = k // 9
            w_nk = w_kn.t().contiguous()                                        # [n_out, 9 * C]
            hit = L.split_f16x2(w_nk.reshape(n_out * 9, c)).reshape(n_out, 9 * 2 * c)
            self._wsplit[key] = hit
        return hit

    @staticmethod
    def _split_act(x):
        n, h, w, c = x.shape
        return L.split_f16x2(x.reshape(n * h * w, c)).reshape(n, h, w, 2 * c)

    def _conv_fw(self, cw, a, residual=None, stride=1, upsample=False):
        if self._tc_ok(cw, stride, upsample):
            a_split = (L.groupnorm(a, None, None, swish=False, out_dtype=torch.float16, normalize=False, upsample=True) if upsample
                       else self._split_act(a))
            return L.tc_conv(a_split, self._split_weight(("fw", id(cw)), cw.w_kn, cw.cout), cw.bias, residual=residual)
        return self.model._conv(cw, a, residual=residual, stride=stride, upsample=upsample, stats=False)

    def _conv_bw(self, name, cw, a, dy, stride=1, upsample=False, need_dx=True):
        """a: the conv's input (NHWC f32); dy: gradient of its output.  Accumulates dW, db; returns dx (or None)."""
        P = self.P
        pad = ((1, 1) if stride == 1 else (0, 0)) if cw.k == 3 else (0, 0)
        if self.use_tc and upsample and L.conv_wgrad_tc_ok(dy, dy, cw.k, stride, False) and cw.cin % 128 == 0:
            a_up = L.groupnorm(a, None, None, swish=False, out_dtype=torch.float32, normalize=False, upsample=True)
            L.conv_wgrad_tc(a_up, dy, P[name + ".weight"].grad)
        elif self.use_tc and L.conv_wgrad_tc_ok(a, dy, cw.k, stride, upsample):
            L.conv_wgrad_tc(a, dy, P[name + ".weight"].grad)             # exact split-fp16 GEMMs over the pixel axis (K = pixels)
        else:
            L.conv_wgrad(a, dy, P[name + ".weight"].grad, kh=cw.k, stride=stride, pad=pad, upsample=upsample)
        L.col_sums(dy.reshape(-1, cw.cout), P[name + ".bias"].grad)
        self._grad_ready(P[name + ".bias"]); self._grad_ready(P[name + ".weight"])
        if not need_dx:
            return None
        wk = cw.w_kn.reshape(cw.k, cw.k, cw.cin, cw.cout)
        if stride == 2:                                         # Downsample: gather form, taps not flipped
            wd = wk.permute(0, 1, 3, 2).reshape(cw.k * cw.k * cw.cout, cw.cin).contiguous()
            return L.simt_conv_dgrad_s2(dy, wd, (a.shape[1], a.shape[2]))
        if self._tc_ok(cw, stride, upsample):
            key = ("bw", id(cw))
            if key not in self._wsplit:                         # a data gradient is a conv with flipped taps and swapped channel roles
                wd = wk.flip(0, 1).permute(0, 1, 3, 2).reshape(cw.k * cw.k * cw.cout, cw.cin).contiguous()
                self._split_weight(key, wd, cw.cin)
            dx = L.tc_conv(self._split_act(dy), self._wsplit[key], None)
            return L.sumpool2x2(dx) if upsample else dx
        wd = wk.flip(0, 1).permute(0, 1, 3, 2).reshape(cw.k * cw.k * cw.cout, cw.cin).contiguous()      # a data gradient is a conv with flipped taps
        dx = L.simt_conv(dy, wd, None, kh=cw.k, stride=1, pad=(1, 1) if cw.k == 3 else (0, 0))
        return L.sumpool2x2(dx) if upsample else dx

    def _lin_bw(self, name, ln, x_rows, dy_rows, residual=None):
        """y = x W^T + b.  Accumulates dW [out,in], db; returns dx = dy W (+ residual)."""
        P = self.P
        m = x_rows.shape[0]
        L.conv_wgrad(x_rows.reshape(1, m, 1, ln.k), dy_rows.reshape(1, m, 1, ln.n), P[name + ".weight"].grad, kh=1, pad=(0, 0), so=(1, ln.k))
        L.col_sums(dy_rows, P[name + ".bias"].grad)
        self._grad_ready(P[name + ".bias"]); self._grad_ready(P[name + ".weight"])
        dx = torch.empty((m, ln.k), dtype=torch.float32, device=x_rows.device)
        L.simt_gemm(dy_rows, ln.w, dx, M=m, N=ln.k, K=ln.n, a_strides=(ln.n, 1), b_strides=(ln.k, 1), ldc=ln.k, residual=residual)
        return dx

    # ------------------------------------------------------------------ blocks
    def _res_fw(self, r, x, tape, name):
        ex = self.model.exact
        st1 = L.gn_mean_rstd(x)
        h = self._conv_fw(r["c1"], self._gn_apply(x, st1, r["n1"], True))
        st2 = L.gn_mean_rstd(h)
        a2 = self._gn_apply(h, st2, r["n2"], True)
        n, hh, ww, c = x.shape
        res = linear(ex, x.reshape(-1, c), r["sc"], torch.float32).reshape(n, hh, ww, -1) if "sc" in r else x
        y = self._conv_fw(r["c2"], a2, residual=res)
        tape.append(("res", name, r, x, st1, h, st2))
        return y

    def _res_bw(self, entry, dy):
        _, name, r, x, st1, h, st2 = entry
        P = self.P
        a2 = self._gn_apply(h, st2, r["n2"], True)
        da2 = self._conv_bw(name + ".conv2", r["c2"], a2, dy)
        dh = L.groupnorm_bwd(h, da2, st2, r["n2"][0], r["n2"][1], P[name + ".norm2.weight"].grad, P[name + ".norm2.bias"].grad, swish=True)
        self._grad_ready(P[name + ".norm2.bias"]); self._grad_ready(P[name + ".norm2.weight"])
        a1 = self._gn_apply(x, st1, r["n1"], True)
        da1 = self._conv_bw(name + ".conv1", r["c1"], a1, dh)
        n, hh, ww, c = x.shape
        if "sc" in r:
            dres = self._lin_bw(name + ".nin_shortcut", r["sc"], x.reshape(-1, c), dy.reshape(-1, dy.shape[-1])).reshape(x.shape)
        else:
            dres = dy
        dx = L.groupnorm_bwd(x, da1, st1, r["n1"][0], r["n1"][1], P[name + ".norm1.weight"].grad, P[name + ".norm1.bias"].grad, swish=True, add=dres)
        self._grad_ready(P[name + ".norm1.bias"]); self._grad_ready(P[name + ".norm1.weight"])
        return dx

    def _attn_fw(self, aw, x, tape, name):
        ex = self.model.exact
        n, hh, ww, c = x.shape
        hw = hh * ww
        st = L.gn_mean_rstd(x)
        a = self._gn_apply(x, st, aw["norm"], False).reshape(n * hw, c)
        qk = linear(ex, a, aw["qk"], torch.float32)                                   # [rows, 2c] = q | k
        v = linear(ex, a, aw["v"], torch.float32)                                     # [rows, c]
        scale = float(int(c) ** (-0.5))
        S = torch.empty((n, hw, hw), dtype=torch.float32, device=x.device)
        L.simt_gemm(qk, qk, S, M=hw, N=hw, K=c, a_strides=(2 * c, 1), b_strides=(1, 2 * c), ldc=hw, batch=(n, 1), a_bs=(hw * 2 * c, 0),
                    b_bs=(hw * 2 * c, 0), c_bs=(hw * hw, 0), b_off=c, alpha=scale)
        Pm = torch.empty_like(S)
        L.softmax_rows(S, Pm, rows_total=n * hw, rows_per_batch=hw, cols=hw, ld_in=hw, ld_out=hw)
        o = torch.empty((n * hw, c), dtype=torch.float32, device=x.device)
        L.simt_gemm(Pm, v, o, M=hw, N=c, K=hw, a_strides=(hw, 1), b_strides=(c, 1), ldc=c, batch=(n, 1), a_bs=(hw * hw, 0), b_bs=(hw * c, 0),
                    c_bs=(hw * c, 0))
        y = linear(ex, o, aw["proj"], torch.float32, residual=x.reshape(n * hw, c)).reshape(x.shape)
        tape.append(("attn", name, aw, x, st, qk, v, Pm, o))
        return y

    def _attn_bw(self, entry, dy):
        _, name, aw, x, st, qk, v, Pm, o = entry
        P = self.P
        n, hh, ww, c = x.shape
        hw = hh * ww
        scale = float(int(c) ** (-0.5))
        dyr = dy.reshape(n * hw, c)
        do = self._lin_bw(name + ".proj_out", aw["proj"], o, dyr)
        dP = torch.empty_like(Pm)                                                      # dP = do v^T
        L.simt_gemm(do, v, dP, M=hw, N=hw, K=c, a_strides=(c, 1), b_strides=(1, c), ldc=hw, batch=(n, 1), a_bs=(hw * c, 0), b_bs=(hw * c, 0),
                    c_bs=(hw * hw, 0))
        dv = torch.empty_like(v)                                                       # dv = P^T do
        L.simt_gemm(Pm, do, dv, M=hw, N=c, K=hw, a_strides=(1, hw), b_strides=(c, 1), ldc=c, batch=(n, 1), a_bs=(hw * hw, 0), b_bs=(hw * c, 0),
                    c_bs=(hw * c, 0))
        dS = L.softmax_bwd_rows(Pm, dP)
        dqk = torch.empty_like(qk)
        L.simt_gemm(dS, qk, dqk, M=hw, N=c, K=hw, a_strides=(hw, 1), b_strides=(2 * c, 1), ldc=2 * c, batch=(n, 1), a_bs=(hw * hw, 0),
                    b_bs=(hw * 2 * c, 0), c_bs=(hw * 2 * c, 0), b_off=c, alpha=scale)                       # dq = scale dS k
        L.simt_gemm(dS, qk, dqk, M=hw, N=c, K=hw, a_strides=(1, hw), b_strides=(2 * c, 1), ldc=2 * c, batch=(n, 1), a_bs=(hw * hw, 0),
                    b_bs=(hw * 2 * c, 0), c_bs=(hw * 2 * c, 0), c_off=c, alpha=scale)                       # dk = scale dS^T q
        a = self._gn_apply(x, st, aw["norm"], False).reshape(n * hw, c)
        da = self._lin_bw(name + ".v", aw["v"], a, dv)
        da = self._lin_bw(name + ".qk", aw["qk"], a, dqk, residual=da)
        dx = L.groupnorm_bwd(x, da.reshape(x.shape), st, aw["norm"][0], aw["norm"][1], P[name + ".norm.weight"].grad, P[name + ".norm.bias"].grad,
                             swish=False, add=dy)
        self._grad_ready(P[name + ".norm.bias"]); self._grad_ready(P[name + ".norm.weight"])
        return dx

    # ------------------------------------------------------------------ the step
    def forward_backward(self, x_nchw):
        """x f32 NCHW in [-1,1] -> loss (python float).  Leaves the gradient (summed over ranks once the handles complete) in flat_g."""
        model, cfg, w = self.model, self.cfg, self.model._w
        self.P = self._pmap()
        self.flat_g.zero_()
        self._handles, self.launched = [], []
        self._bucket_left = list(self._bucket_size)
        was_training = model.training
        model.training = True                                      # QuantizeEMA.forward: EMA statistics + codebook overwrite (utils_th.py:46-64)
        x = L.nchw_to_nhwc(model._in(x_nchw))
        tape = []
        e, d = w["enc"], w["dec"]
        # ---------------- encoder
        h = self._conv_fw(e["conv_in"], x)
        tape.append(("conv", "encoder.conv_in", e["conv_in"], x, 1, False, False))
        for lv, lvw in enumerate(e["levels"]):
            for b, r in enumerate(lvw["blocks"]):
                h = self._res_fw(r, h, tape, f"encoder.down.{lv}.block.{b}")
                if lvw["attns"]:
                    h = self._attn_fw(lvw["attns"][b], h, tape, f"encoder.down.{lv}.attn.{b}")
            if lvw["down"] is not None:
                tape.append(("conv", f"encoder.down.{lv}.downsample.conv", lvw["down"], h, 2, False, True))
                h = self._conv_fw(lvw["down"], h, stride=2)
        h = self._res_fw(e["mid1"], h, tape, "encoder.mid.block_1")
        h = self._attn_fw(e["mida"], h, tape, "encoder.mid.attn_1")
        h = self._res_fw(e["mid2"], h, tape, "encoder.mid.block_2")
        st = L.gn_mean_rstd(h)
        a = self._gn_apply(h, st, e["norm_out"], True)
        tape.append(("normconv", "encoder.norm_out", "encoder.conv_out", e, h, st))
        hz = self._conv_fw(e["conv_out"], a)
        n, zh, zw, zc = hz.shape
        # ---------------- quantizer (utils_th.py:32-68): z rows, nearest code, straight-through
        z = linear(model.exact, hz.reshape(-1, zc), w["quant_conv"], torch.float32)
        quant, diff, idx = model._quantize(z, want_quant=True)      # training: EMA update + packed all-reduce inside
        pq = linear(model.exact, quant, w["post_quant_conv"], torch.float32).reshape(n, zh, zw, -1)
        # ---------------- decoder
        g = self._conv_fw(d["conv_in"], pq)
        tape.append(("conv", "decoder.conv_in", d["conv_in"], pq, 1, False, True))
        g = self._res_fw(d["mid1"], g, tape, "decoder.mid.block_1")
        g = self._attn_fw(d["mida"], g, tape, "decoder.mid.attn_1")
        g = self._res_fw(d["mid2"], g, tape, "decoder.mid.block_2")
        for lv in reversed(range(len(cfg.ch_mult))):
            lvw = d["levels"][lv]
            for b, r in enumerate(lvw["blocks"]):
                g = self._res_fw(r, g, tape, f"decoder.up.{lv}.block.{b}")
                if lvw["attns"]:
                    g = self._attn_fw(lvw["attns"][b], g, tape, f"decoder.up.{lv}.attn.{b}")
            if lvw["up"] is not None:
                tape.append(("conv", f"decoder.up.{lv}.upsample.conv", lvw["up"], g, 1, True, True))
                g = self._conv_fw(lvw["up"], g, upsample=True)
        st = L.gn_mean_rstd(g)
        a = self._gn_apply(g, st, d["norm_out"], True)
        tape.append(("normconv", "decoder.norm_out", "decoder.conv_out", d, g, st))
        dec = self._conv_fw(d["conv_out"], a)
        # ---------------- loss (vqgan_th.py:400-411): mean |x - xrec| + codebook_weight * diff
        ddec, l1 = L.l1_grad(x, dec, 1.0 / dec.numel())
        rec = l1 / dec.numel()
        loss = rec.to(torch.float32).reshape(()) + float(cfg.codebook_weight) * diff
        self.last = dict(rec_loss=rec, quant_loss=diff, codes=idx.reshape(n, zh, zw), reconstruction=dec)
        # ---------------- backward
        dy = ddec
        P = self.P
        for entry in reversed(tape):
            kind = entry[0]
            if kind == "res":
                dy = self._res_bw(entry, dy)
            elif kind == "attn":
                dy = self._attn_bw(entry, dy)
            elif kind == "conv":
                _, name, cw, xin, stride, ups, need_dx = entry
                dy = self._conv_bw(name, cw, xin, dy, stride=stride, upsample=ups, need_dx=need_dx)
                if name == "decoder.conv_in":
                    # through post_quant_conv, the straight-through estimator and the commitment term, quant_conv
                    dq = self._lin_bw("post_quant_conv", w["post_quant_conv"], quant, dy.reshape(-1, dy.shape[-1]))
                    dz = L.lincomb3(1.0, dq, 2.0 * float(cfg.codebook_weight) / z.numel(), z, -2.0 * float(cfg.codebook_weight) / z.numel(), quant)
                    if model.quantizer == "commit":
                        # d/dE of beta mean((q - sg(z))^2): column k gets 2 beta / numel * (count_k e_k - sum of the z rows mapped to k);
                        # the straight-through output carries no gradient to E (utils_th.py:117)
                        pe = P["quantize.embeddings"]
                        counts, zsum = L.vq_ema_stats(z, idx, pe.tensor.shape[1])
                        L.vq_commit_grad(pe.tensor, counts, zsum, 2.0 * model.beta * float(cfg.codebook_weight) / z.numel(), pe.grad)
                        self._grad_ready(pe)
                    dy = self._lin_bw("quant_conv", w["quant_conv"], hz.reshape(-1, zc), dz).reshape(hz.shape)
            elif kind == "normconv":
                _, nname, cname, blk, xin, st = entry
                nw = blk["norm_out"]
                cw = blk["conv_out"]
                a = self._gn_apply(xin, st, nw, True)
                da = self._conv_bw(cname, cw, a, dy)
                dy = L.groupnorm_bwd(xin, da, st, nw[0], nw[1], P[nname + ".weight"].grad, P[nname + ".bias"].grad, swish=True)
                self._grad_ready(P[nname + ".bias"]); self._grad_ready(P[nname + ".weight"])
        model.training = was_training
        if any(self._bucket_left):
            raise RuntimeError("backward pass left gradient buckets incomplete: " + str([self.buckets[i][2] for i, n in enumerate(self._bucket_left) if n]))
        return loss

    def optimizer_step(self):
        for h in self._handles:
            h.wait()
        self._handles = []
        self.step_count += 1
        L.adam(self.flat_p, self.flat_g, self.flat_m, self.flat_v, lr=self.lr, beta1=self.betas[0], beta2=self.betas[1], eps=self.eps,
               step=self.step_count, grad_scale=1.0 / self._world())
        self._wsplit = {}                               # split-fp16 operand copies of the conv weights are stale now
        if self.model.quantizer == "commit":
            self.model._refresh_codebook()
        else:
            self.model._refresh_decode_table()

    def training_step(self, batch, batch_idx=0):
        """vqgan_th.py:413-423 + the optimizer step Lightning runs after it.  Returns the loss of the step (0-d f32 tensor)."""
        loss = self.forward_backward(batch)
        self.optimizer_step()
        return loss

    # ------------------------------------------------------------------ export (reference layouts)
    def _export(self, get):
        out = {}
        for p in self.params:
            t = get(p)
            if p.kind == "conv":
                cout = t.shape[1]
                k = int(round(math.sqrt(t.shape[0] // p.cin)))
                out[p.name] = t.reshape(k, k, -1, cout).permute(3, 2, 0, 1).contiguous().cpu()
            elif p.kind == "dense":
                t4 = t.reshape(t.shape[0], t.shape[1], 1, 1).cpu()
                if p.part:
                    half = t.shape[0] // 2
                    out[p.part[0] + ".weight"], out[p.part[1] + ".weight"] = t4[:half].clone(), t4[half:].clone()
                else:
                    out[p.name] = t4.clone()
            else:
                tc = t.cpu().clone()
                if p.part:
                    half = tc.shape[0] // 2
                    out[p.part[0] + ".bias"], out[p.part[1] + ".bias"] = tc[:half].clone(), tc[half:].clone()
                else:
                    out[p.name] = tc
        return out

    def export_gradients(self):
        """Gradient of the last forward_backward (already summed over ranks if the handles were waited for), reference layouts."""
        return self._export(lambda p: p.grad)

    def export_state_dict(self):
        """Reference-keyed state_dict after training; also refreshes the model's host copy."""
        sd = self.model.state_dict()
        sd.update(self._export(lambda p: p.tensor))
        self.model._sd = {k: v.clone() for k, v in sd.items()}
        return sd
