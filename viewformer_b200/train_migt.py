"""Transformer training step — MIGT.train_step (viewformer/models/migt.py:464-505) on libvf_b200 kernels, fp32.

    forward   three streams as in ``call(compute_losses=True, training=True)`` (migt.py:338-455): tokens + poses, MASK tokens + query
              poses (image generation), tokens + LOC token (localisation); block-causal multi-end attention
              (branching_attention.py:82-126); dropout at the reference's four sites (embeddings, attention weights, attention
              output, MLP output) from a stateless hash generator
    loss      mean over the batch of  image_generation_weight * CE(stream 1 logits, tokens)[views >= n_loss_skip]
              + localization_weight * (position MSE + orientation MSE of the stream-2 pose head)
    backward  hand-written: dense layers on the exact split-fp16 tensor-core GEMM (forward, data and weight gradient; VF_TRAIN_TC=0: the
              fp32 CUDA-core kernels), attention through vf_simt_gemm (strided, batched per (scene, head)), vf_conv_wgrad for the remaining weight
              gradients, LayerNorm / GELU / softmax / embedding / loss kernels of vf_backward.cu
    update    per-tensor tf.clip_by_norm when gradient_clip_val > 0 (migt.py:486-487), AdamWeightDecay = Keras Adam preceded by the
              decoupled decay lr*wd*p for every variable whose name has no "bias" (models/utils.py:424, 507-515 — LayerNorm gamma /
              beta ARE decayed: their Keras names are ln_1/gamma ..., which match none of the exclusion patterns), learning rate
              = 2000-step linear warm-up then cosine decay to 0 (migt.py:457-462, models/utils.py:310-416; the first step runs at lr 0)
    exchange  gradients live in one flat buffer ordered by backward completion; contiguous buckets are all-reduced (SUM, then
              divided by the world size — see the note in DESIGN.md on MirroredStrategy's per-replica reduce_mean) asynchronously
              while the rest of the backward pass runs.

Parameters are kept in the reference's own layouts (Conv1D weight [in, out], bias [1, out]); ``state_dict()`` can be loaded straight
into ``viewformer_b200.MIGT`` for inference.
"""
import math
import os
from collections import OrderedDict

import torch

from . import _lib as L

LN_EPS = 1e-5


class MIGTTrainer:
    def __init__(self, model, betas=(0.9, 0.999), eps=1e-8, warmup_steps=2000, bucket_bytes=64 << 20, process_group=None, seed=0,
                 grad_reduce="sum"):
        cfg = model.config
        if cfg.random_pose_multiplier != 1.0:
            raise NotImplementedError("random_pose_multiplier != 1 (pose-scale augmentation, migt.py:350-353) is not supported")
        self.model, self.cfg, self.device = model, cfg, model.device
        self.betas, self.eps, self.warmup_steps = betas, eps, warmup_steps
        self.group, self.bucket_bytes, self.seed = process_group, bucket_bytes, seed
        # "sum": what the reference does — every replica takes tf.reduce_mean of ITS loss and MirroredStrategy sums the replica
        # gradients (migt.py:471-476: "the learning rate should be scaled accordingly"); "mean" divides by the world size instead
        assert grad_reduce in ("sum", "mean")
        self.grad_reduce = grad_reduce
        self.iterations = 0                                        # optimizer.iterations (0-based: the schedule sees it BEFORE the increment)
        self.use_tc = os.environ.get("VF_TRAIN_TC", "1") != "0"
        self._wsplit = {}                                          # split-fp16 operand copies of the weights, rebuilt after every step
        self.use_loc = model.use_localization
        from .schedules import parse
        self._loc_schedule = parse(cfg.localization_weight).with_total_steps(int(cfg.total_steps))
        self.dynamic_pose = bool(cfg.use_dynamic_pose_loss) and self.use_loc
        self._build(model.state_dict())

    @property
    def loc_weight(self):
        """localization_weight(self._train_counter) (migt.py:446): the schedule at the number of steps taken so far."""
        return float(self._loc_schedule(self.iterations)) if self.use_loc else 0.0

    # ------------------------------------------------------------------ parameters
    def _build(self, sd):
        names = list(self.model.param_shapes().keys())
        # backward completion order: heads, ln_f, blocks from last to first, pose embedding, wpe, wte (tied: complete only at the very end)
        n_layer = self.cfg.n_layer
        order = ([k for k in names if k.startswith("pose_loss_weighting_criterion.")] + [k for k in names if k.startswith("pose_classifier.")]
                 + [k for k in names if k.startswith("ln_f.")])
        for i in reversed(range(n_layer)):
            order += [k for k in names if k.startswith(f"h.{i}.")]
        order += [k for k in names if k.startswith("pose_embedding.")] + ["wpe.embeddings", "wte.weight"]
        assert sorted(order) == sorted(names)
        offs, n = {}, 0
        for k in order:
            offs[k] = n
            n += (sd[k].numel() + 3) // 4 * 4
        dev = self.device
        self.flat_p = torch.zeros((n,), dtype=torch.float32, device=dev)
        self.flat_g, self.flat_m, self.flat_v = torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p), torch.zeros_like(self.flat_p)
        self.p, self.g = OrderedDict(), OrderedDict()
        for k in order:
            t = sd[k].to(torch.float32)
            self.p[k] = self.flat_p[offs[k]:offs[k] + t.numel()].view(t.shape)
            self.p[k].copy_(t)
            self.g[k] = self.flat_g[offs[k]:offs[k] + t.numel()].view(t.shape)
        self.order, self.offs = order, offs
        self.buckets, start = [], 0
        for i, k in enumerate(order):
            end = offs[k] + (sd[k].numel() + 3) // 4 * 4
            if (end - start) * 4 >= self.bucket_bytes or i == len(order) - 1:
                self.buckets.append((start, end))
                start = end
        self._bucket_of, bi = {}, 0
        for k in order:
            while offs[k] >= self.buckets[bi][1]:
                bi += 1
            self._bucket_of[k] = bi
        self._bucket_size = [sum(1 for b in self._bucket_of.values() if b == i) for i in range(len(self.buckets))]
        # AdamWeightDecay: every variable except the ones whose name contains "bias" (see the module docstring)
        self.decay = {k: ("bias" not in k) for k in order}

    def state_dict(self):
        return OrderedDict((k, self.p[k].detach().cpu().clone()) for k in self.model.param_shapes().keys())

    def gradients(self):
        return OrderedDict((k, self.g[k].detach().cpu().clone()) for k in self.model.param_shapes().keys())

    # ------------------------------------------------------------------ exchange
    def _world(self):
        import torch.distributed as dist
        return dist.get_world_size(self.group) if (dist.is_available() and dist.is_initialized()) else 1

    def _ready(self, *names):
        for k in names:
            b = self._bucket_of[k]
            self._left[b] -= 1
            if self._left[b] == 0:
                self.launched.append(b)
                if self._world() > 1:
                    import torch.distributed as dist
                    s, e = self.buckets[b]
                    self._handles.append(dist.all_reduce(self.flat_g[s:e], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            elif self._left[b] < 0:
                raise RuntimeError(f"gradient of {k} signalled twice")

    # ------------------------------------------------------------------ dense layer (Conv1D: x @ W[in,out] + b[1,out])
    # Dense layers whose sizes fit the tensor-core tiles run forward, data gradient and weight gradient on the exact split-fp16 GEMM
    # (fp32-faithful: three fp16 MMA passes, chunked accumulation — DESIGN.md 5.3); VF_TRAIN_TC=0 keeps everything on the CUDA cores.
    def _tc_dense_ok(self, k, n):
        return self.use_tc and k % 128 == 0 and n % 128 == 0

    def _wsplit_get(self, key, make):
        hit = self._wsplit.get(key)
        if hit is None:
            hit = self._wsplit[key] = L.split_f16x2(make())
        return hit

    def _dense_fw_tc(self, x, W_nk_key, W_nk_make, n, k, bias=None, residual=None):
        """x [M, k] fp32 @ W^T with W given K-major ([n, k]) -> [M, n] fp32."""
        out = torch.empty((x.shape[0], n), dtype=torch.float32, device=x.device)
        L.tc_gemm(L.split_f16x2(x), self._wsplit_get(W_nk_key, W_nk_make), out, M=x.shape[0], N=n, K=k, lda=2 * k, ldb=2 * k, ldc=n,
                  bias=bias, bias_mode=L.BIAS_N if bias is not None else L.BIAS_NONE, residual=residual, lo_a=k, lo_b=k)
        return out

    def _lin(self, x, name, act=L.ACT_NONE, residual=None):
        W, b = self.p[name + ".weight"], self.p[name + ".bias"]
        k, n = W.shape
        if act == L.ACT_NONE and self._tc_dense_ok(k, n):
            return self._dense_fw_tc(x, ("fw", name), lambda: W.t().contiguous(), n, k, bias=b.reshape(-1), residual=residual)
        out = torch.empty((x.shape[0], n), dtype=torch.float32, device=x.device)
        L.simt_gemm(x, W, out, M=x.shape[0], N=n, K=k, a_strides=(k, 1), b_strides=(n, 1), ldc=n, bias=b.reshape(-1), bias_mode=L.BIAS_N,
                    act=act, residual=residual)
        return out

    def _lin_bw(self, x, dy, name, need_dx=True, residual=None):
        W = self.p[name + ".weight"]
        k, n = W.shape
        m = x.shape[0]
        tc = self._tc_dense_ok(k, n)
        if tc:
            L.dense_wgrad_tc(x, dy, self.g[name + ".weight"])
        else:
            L.conv_wgrad(x.reshape(1, m, 1, k), dy.reshape(1, m, 1, n), self.g[name + ".weight"], kh=1, pad=(0, 0), so=(n, 1))
        L.col_sums(dy, self.g[name + ".bias"].reshape(-1))
        self._ready(name + ".bias", name + ".weight")
        if not need_dx:
            return None
        if tc:                                  # dx = dy W^T: W [k, n] is already K-major for this product
            return self._dense_fw_tc(dy, ("bw", name), lambda: W.contiguous(), k, n, residual=residual)
        dx = torch.empty((m, k), dtype=torch.float32, device=x.device)
        L.simt_gemm(dy, W, dx, M=m, N=k, K=n, a_strides=(n, 1), b_strides=(1, n), ldc=k, residual=residual)
        return dx

    def _ln(self, x, name):
        return L.layernorm(x, self.p[name + ".gamma"], self.p[name + ".beta"], torch.float32, eps=LN_EPS)

    def _ln_bw(self, x, dy, name, add=None, last=True):
        dx = L.layernorm_bwd(x, dy, self.p[name + ".gamma"], self.g[name + ".gamma"], self.g[name + ".beta"], eps=LN_EPS, add=add)
        if last:
            self._ready(name + ".beta", name + ".gamma")
        return dx

    def _drop(self, x, site):
        rate = float(self.cfg.dropout)
        if rate <= 0.0:
            return x
        return L.dropout(x, rate, (self.seed * 1000003 + self.iterations) * 4096 + site)

    # ------------------------------------------------------------------ attention over the stream list
    def _attention_fw(self, vqk, B, S, Lt, site0):
        """vqk: per stream [B*S, 3d] = v | q | k.  Returns (outputs [B*S, d] per stream, saved probabilities)."""
        d, H = self.cfg.d_model, self.cfg.n_head
        dh = d // H
        dev = vqk[0].device
        outs, probs = [], []
        for s, t in enumerate(vqk):
            cols = S if s == 0 else 2 * S
            sc = torch.empty((B, H, S, cols), dtype=torch.float32, device=dev)
            for half, ks in ((0, 0),) if s == 0 else ((0, 0), (1, s)):
                L.simt_gemm(t, vqk[ks], sc, M=S, N=S, K=dh, a_strides=(3 * d, 1), b_strides=(1, 3 * d), ldc=cols, batch=(B, H),
                            a_bs=(S * 3 * d, dh), b_bs=(S * 3 * d, dh), c_bs=(H * S * cols, S * cols), a_off=d, b_off=2 * d, c_off=half * S)
            P = torch.empty_like(sc)
            L.softmax_rows(sc, P, rows_total=B * H * S, rows_per_batch=S, cols=cols, ld_in=cols, ld_out=cols, mask_mode=1 if s == 0 else 2, block=Lt)
            Pd = self._drop(P, site0 + s)
            o = torch.empty((B * S, d), dtype=torch.float32, device=dev)
            L.simt_gemm(Pd, vqk[0], o, M=S, N=dh, K=S, a_strides=(cols, 1), b_strides=(3 * d, 1), ldc=d, batch=(B, H), a_bs=(H * S * cols, S * cols),
                        b_bs=(S * 3 * d, dh), c_bs=(S * d, dh))
            if s > 0:
                L.simt_gemm(Pd, t, o, M=S, N=dh, K=S, a_strides=(cols, 1), b_strides=(3 * d, 1), ldc=d, batch=(B, H), a_bs=(H * S * cols, S * cols),
                            b_bs=(S * 3 * d, dh), c_bs=(S * d, dh), a_off=S, residual=o)
            outs.append(o)
            probs.append(P)
        return outs, probs

    def _attention_bw(self, vqk, probs, dos, B, S, Lt, site0):
        d, H = self.cfg.d_model, self.cfg.n_head
        dh = d // H
        dev = vqk[0].device
        dvqk = [torch.zeros_like(t) for t in vqk]
        for s, (t, P, do) in enumerate(zip(vqk, probs, dos)):
            cols = S if s == 0 else 2 * S
            Pd = self._drop(P, site0 + s)
            pb = (H * S * cols, S * cols)
            dP = torch.empty_like(P)
            for half, ks in ((0, 0),) if s == 0 else ((0, 0), (1, s)):
                # dP[:, half] = do v_ks^T ;  dv_ks += Pd[:, half]^T do
                L.simt_gemm(do, vqk[ks], dP, M=S, N=S, K=dh, a_strides=(d, 1), b_strides=(1, 3 * d), ldc=cols, batch=(B, H), a_bs=(S * d, dh),
                            b_bs=(S * 3 * d, dh), c_bs=pb, c_off=half * S)
                L.simt_gemm(Pd, do, dvqk[ks], M=S, N=dh, K=S, a_strides=(1, cols), b_strides=(d, 1), ldc=3 * d, batch=(B, H), a_bs=pb, b_bs=(S * d, dh),
                            c_bs=(S * 3 * d, dh), a_off=half * S, residual=dvqk[ks])
            if float(self.cfg.dropout) > 0:
                dP = self._drop(dP, site0 + s)                      # same mask and scale as the forward pass
            dS = L.softmax_bwd_rows(P, dP)
            for half, ks in ((0, 0),) if s == 0 else ((0, 0), (1, s)):
                # dq_s += dS[:, half] k_ks ;  dk_ks += dS[:, half]^T q_s
                L.simt_gemm(dS, vqk[ks], dvqk[s], M=S, N=dh, K=S, a_strides=(cols, 1), b_strides=(3 * d, 1), ldc=3 * d, batch=(B, H), a_bs=pb,
                            b_bs=(S * 3 * d, dh), c_bs=(S * 3 * d, dh), a_off=half * S, b_off=2 * d, c_off=d, residual=dvqk[s])
                L.simt_gemm(dS, t, dvqk[ks], M=S, N=dh, K=S, a_strides=(1, cols), b_strides=(3 * d, 1), ldc=3 * d, batch=(B, H), a_bs=pb,
                            b_bs=(S * 3 * d, dh), c_bs=(S * 3 * d, dh), a_off=half * S, b_off=d, c_off=2 * d, residual=dvqk[ks])
        return dvqk

    # ------------------------------------------------------------------ the step
    def forward_backward(self, poses, tokens):
        cfg, dev, p, g = self.cfg, self.device, self.p, self.g
        self.flat_g.zero_()
        self._handles, self.launched, self._left = [], [], list(self._bucket_size)
        tokens = torch.as_tensor(tokens)
        B, T = tokens.shape[:2]
        Lt, d, V = self.model.n_image_tokens, cfg.d_model, cfg.n_embeddings
        S, skip = T * Lt, cfg.n_loss_skip
        ids = tokens.reshape(B, T, Lt).to(device=dev, dtype=torch.int32).contiguous()
        poses = torch.as_tensor(poses, dtype=torch.float32).to(dev).reshape(B * T, 7).contiguous()
        mult = torch.tensor([cfg.pose_multiplier] * 3 + [1.0] * 4, dtype=torch.float32, device=dev)
        pin = (poses * mult).contiguous()                                           # get_model_input (migt.py:139-145)
        # ---------------- embeddings: three streams (migt.py:354-405)
        pe_h = self._lin(pin, "pose_embedding.c_fc")
        pe = self._lin(self._gelu(pe_h), "pose_embedding.c_proj")                # [B*T, d]
        wte, wpe = p["wte.weight"], p["wpe.embeddings"]
        loc_rows = wte[self.model.localization_token].reshape(1, d).expand(B * T, d).contiguous()
        xs = [L.migt_embed(ids, 0, wte, wpe, pe, B * T, Lt), L.migt_embed(None, self.model.mask_token, wte, wpe, pe, B * T, Lt)]
        if self.use_loc:
            xs.append(L.migt_embed(ids, 0, wte, wpe, loc_rows, B * T, Lt))
        ns = len(xs)
        xs = [self._drop(x, 10 + s) for s, x in enumerate(xs)]
        tape = []
        for li in range(cfg.n_layer):
            pre = f"h.{li}."
            site = 100 + li * 20
            a = [self._ln(x, pre + "ln_1") for x in xs]
            vqk = [self._lin(t, pre + "attn.c_attn") for t in a]
            outs, probs = self._attention_fw(vqk, B, S, Lt, site)
            ys = [self._drop(self._lin(o, pre + "attn.c_proj"), site + 4 + s) for s, o in enumerate(outs)]
            ys = [L.lincomb3(1.0, x, 1.0, y) for x, y in zip(xs, ys)]
            hm = [self._lin(self._ln(y, pre + "ln_2"), pre + "mlp.c_fc") for y in ys]
            zs = [self._drop(self._lin(self._gelu(h), pre + "mlp.c_proj"), site + 8 + s) for s, h in enumerate(hm)]
            zs = [L.lincomb3(1.0, y, 1.0, z) for y, z in zip(ys, zs)]
            tape.append((xs, vqk, probs, outs, ys, hm))
            xs = zs
        # ---------------- heads and losses (migt.py:408-452)
        hn = [self._ln(x, "ln_f") for x in xs]
        denom = float(B * (T - skip) * Lt)
        view_ok = (torch.arange(T, device=dev) >= skip).to(torch.float32).repeat_interleave(Lt).repeat(B)        # [B*S] row mask
        head_tc = self._tc_dense_ok(d, V)
        if head_tc:                                                                                                # tied head, first V rows (:417)
            logits = self._dense_fw_tc(hn[1], ("fw", "wte"), lambda: wte[:V].contiguous(), V, d)
        else:
            logits = torch.empty((B * S, V), dtype=torch.float32, device=dev)
            L.simt_gemm(hn[1], wte, logits, M=B * S, N=V, K=d, a_strides=(d, 1), b_strides=(1, d), ldc=V)
        ce_rows = L.cross_entropy_rows(logits, ids.reshape(-1), float(cfg.label_smoothing))
        ce = L.row_mean(ce_rows.reshape(B, S), skip * Lt)
        loss = ce * float(cfg.image_generation_weight)
        self.last = dict(ce_loss=ce, logits=logits.reshape(B, T, Lt, V))
        dhn = [None] * ns
        dlog = L.cross_entropy_grad(logits, ids.reshape(-1), (view_ok * (float(cfg.image_generation_weight) / denom)).contiguous(), float(cfg.label_smoothing))
        # tied LM head backward: d hn1 = dlogits wte[:V];  d wte[:V] += dlogits^T hn1
        if head_tc:
            dhn[1] = self._dense_fw_tc(dlog, ("bw", "wte"), lambda: wte[:V].t().contiguous(), d, V)
            L.dense_wgrad_tc(dlog, hn[1], g["wte.weight"][:V])
        else:
            dhn[1] = torch.empty_like(hn[1])
            L.simt_gemm(dlog, wte, dhn[1], M=B * S, N=d, K=V, a_strides=(V, 1), b_strides=(d, 1), ldc=d)
            L.conv_wgrad(dlog.reshape(1, B * S, 1, V), hn[1].reshape(1, B * S, 1, d), g["wte.weight"], kh=1, pad=(0, 0), so=(d, 1))
        if self.use_loc:
            pc_h = self._lin(hn[2], "pose_classifier.c_fc")
            raw = self._lin(self._gelu(pc_h), "pose_classifier.c_proj")            # [B*S, 7]
            pl_rows, ol_rows = L.pose_loss_rows(raw, poses, Lt, float(cfg.pose_multiplier))
            pl, ol = L.row_mean(pl_rows.reshape(B, S), skip * Lt), L.row_mean(ol_rows.reshape(B, S), skip * Lt)
            lw_now = self.loc_weight
            if self.dynamic_pose:
                # DynamicLossWeightingCriterion (migt.py:107-120): P = sum_b (w0 + e^-w0 pos_b) + (w1 + e^-w1 ori_b), a scalar added to every
                # scene's loss; d mean(loss) / d w = lw (B - e^-w sum_b loss_b), d / d pos_b = lw e^-w0 (B times the plain-sum case)
                wkey = "pose_loss_weighting_criterion.pos_ori_weights"
                w01 = self.p[wkey].detach().double().cpu()
                e0, e1 = math.exp(-float(w01[0])), math.exp(-float(w01[1]))
                pls, ols = float(pl.double().sum()), float(ol.double().sum())
                pose_loss = torch.full_like(pl, float(B * (w01[0] + w01[1]) + e0 * pls + e1 * ols))
                self.g[wkey].copy_(torch.tensor([lw_now * (B - e0 * pls), lw_now * (B - e1 * ols)], dtype=torch.float32))
                self._ready(wkey)
                ps, os_ = e0 * B, e1 * B
            else:
                pose_loss, ps, os_ = pl + ol, 1.0, 1.0
            loss = loss + pose_loss * lw_now
            self.last.update(pose_pos_loss=pl, pose_ori_loss=ol, pose_loss=pose_loss)
            draw = L.pose_loss_grad(raw, poses, (view_ok * (lw_now / denom)).contiguous(), Lt, float(cfg.pose_multiplier), ps, os_)
            dg_ = self._lin_bw(self._gelu(pc_h), draw, "pose_classifier.c_proj")
            dhn[2] = self._lin_bw(hn[2], L.gelu_bwd(pc_h, dg_), "pose_classifier.c_fc")
        else:
            self._ready(*[k for k in self.order if k.startswith("pose_classifier.") or k.startswith("pose_loss_weighting_criterion.")])
        # ln_f backward (shared parameters: accumulate over the streams that carry a loss)
        dxs = [torch.zeros_like(xs[0])] + [None] * (ns - 1)
        live = [s for s in range(ns) if dhn[s] is not None]
        for s in live:
            dxs[s] = self._ln_bw(xs[s], dhn[s], "ln_f", last=(s == live[-1]))
        # ---------------- blocks, last to first
        for li in reversed(range(cfg.n_layer)):
            pre = f"h.{li}."
            site = 100 + li * 20
            xin, vqk, probs, outs, ys, hm = tape[li]
            dys = []
            for s in range(ns):
                dz = self._drop(dxs[s], site + 8 + s) if float(cfg.dropout) > 0 else dxs[s]
                last = s == ns - 1
                dgel = self._lin_bw_shared(self._gelu(hm[s]), dz, pre + "mlp.c_proj", last)
                dm = self._lin_bw_shared(self._ln(ys[s], pre + "ln_2"), L.gelu_bwd(hm[s], dgel), pre + "mlp.c_fc", last)
                dys.append(self._ln_bw(ys[s], dm, pre + "ln_2", add=dxs[s], last=last))
            dos = []
            for s in range(ns):
                dy = self._drop(dys[s], site + 4 + s) if float(cfg.dropout) > 0 else dys[s]
                dos.append(self._lin_bw_shared(outs[s], dy, pre + "attn.c_proj", s == ns - 1))
            dvqk = self._attention_bw(vqk, probs, dos, B, S, Lt, site)
            new_dxs = []
            for s in range(ns):
                last = s == ns - 1
                da = self._lin_bw_shared(self._ln(xin[s], pre + "ln_1"), dvqk[s], pre + "attn.c_attn", last)
                new_dxs.append(self._ln_bw(xin[s], da, pre + "ln_1", add=dys[s], last=last))
            dxs = new_dxs
        # ---------------- embeddings backward
        if float(cfg.dropout) > 0:
            dxs = [self._drop(dx, 10 + s) for s, dx in enumerate(dxs)]
        dpe = torch.zeros((B * T, d), dtype=torch.float32, device=dev)
        L.migt_embed_bwd(dxs[0], ids, 0, B * T, Lt, g["wte.weight"], g["wpe.embeddings"], dpe)
        L.migt_embed_bwd(dxs[1], None, self.model.mask_token, B * T, Lt, g["wte.weight"], g["wpe.embeddings"], dpe)
        if self.use_loc:
            dloc = torch.zeros((B * T, d), dtype=torch.float32, device=dev)
            L.migt_embed_bwd(dxs[2], ids, 0, B * T, Lt, g["wte.weight"], g["wpe.embeddings"], dloc)
            L.col_sums(dloc, g["wte.weight"][self.model.localization_token])
        dh_ = self._lin_bw(self._gelu(pe_h), dpe, "pose_embedding.c_proj")
        self._lin_bw(pin, L.gelu_bwd(pe_h, dh_), "pose_embedding.c_fc", need_dx=False)
        self._ready("wpe.embeddings", "wte.weight")
        if any(self._left):
            raise RuntimeError("backward pass left gradient buckets incomplete: " + str([i for i, n in enumerate(self._left) if n]))
        self.last["loss_per_scene"] = loss
        return loss.mean()

    def _gelu(self, x):
        return L.gelu(x)

    def _lin_bw_shared(self, x, dy, name, last):
        """_lin_bw for a layer applied to every stream: the weight-gradient kernels accumulate; readiness is signalled on the last one."""
        W = self.p[name + ".weight"]
        k, n = W.shape
        m = x.shape[0]
        tc = self._tc_dense_ok(k, n)
        if tc:
            L.dense_wgrad_tc(x, dy, self.g[name + ".weight"])
        else:
            L.conv_wgrad(x.reshape(1, m, 1, k), dy.reshape(1, m, 1, n), self.g[name + ".weight"], kh=1, pad=(0, 0), so=(n, 1))
        L.col_sums(dy, self.g[name + ".bias"].reshape(-1))
        if last:
            self._ready(name + ".bias", name + ".weight")
        if tc:
            return self._dense_fw_tc(dy, ("bw", name), lambda: W.contiguous(), k, n)
        dx = torch.empty((m, k), dtype=torch.float32, device=x.device)
        L.simt_gemm(dy, W, dx, M=m, N=k, K=n, a_strides=(n, 1), b_strides=(1, n), ldc=k)
        return dx

    # ------------------------------------------------------------------ schedule + optimizer (models/utils.py:310-564)
    def learning_rate(self, step=None):
        step = self.iterations if step is None else step
        init, warm = float(self.cfg.learning_rate), self.warmup_steps
        if warm and step < warm:
            return init * (step / float(warm))
        decay_steps = max(1, int(self.cfg.total_steps) - warm)
        t = min(max(step - warm, 0), decay_steps) / float(decay_steps)
        return init * 0.5 * (1.0 + math.cos(math.pi * t))

    def optimizer_step(self):
        for h in self._handles:
            h.wait()
        self._handles = []
        lr = self.learning_rate()
        self.iterations += 1
        self._wsplit = {}
        wd = float(self.cfg.weight_decay)
        clip = float(self.cfg.gradient_clip_val or 0.0)
        gs = 1.0 / self._world() if self.grad_reduce == "mean" else 1.0
        for k in self.order:
            cs = 1.0
            if clip > 0:                                                  # tf.clip_by_norm: g * clip / max(|g|, clip), per tensor
                nrm = math.sqrt(float(L.sumsq(self.g[k].reshape(-1)))) * gs
                cs = clip / max(nrm, clip)
            o, n = self.offs[k], self.p[k].numel()
            L.adamw_keras(self.flat_p[o:o + n], self.flat_g[o:o + n], self.flat_m[o:o + n], self.flat_v[o:o + n], lr=lr, beta1=self.betas[0],
                          beta2=self.betas[1], eps=self.eps, weight_decay=wd if (wd > 0 and self.decay[k]) else 0.0, step=self.iterations,
                          grad_scale=gs, clip_scale=cs)

    def train_step(self, batch):
        """(poses [B,T,7], tokens [B,T,h,w]) -> dict(loss, ce_loss, [pose losses], acc, learning_rate) — migt.py:464-505."""
        poses, tokens = batch
        loss = self.forward_backward(poses, tokens)
        lr = self.learning_rate()
        self.optimizer_step()
        out = {k: float(torch.as_tensor(v, dtype=torch.float32).mean()) for k, v in self.last.items() if k.endswith("loss")}
        out["loss"] = float(loss)
        tok = torch.as_tensor(tokens).to(self.device)
        logits = self.last["logits"]
        pred = L.argmax_rows(logits.reshape(-1, logits.shape[-1])).reshape(tok.shape[0], tok.shape[1], -1)
        skip = self.cfg.n_loss_skip
        out["acc"] = float((pred[:, skip:] == tok.reshape(tok.shape[0], tok.shape[1], -1)[:, skip:]).float().mean())
        out["learning_rate"] = lr
        return out
