"""Independent PyTorch (CPU, fp32) restatement of the Pocket-TTS generation path — the "second opinion"
that pins oracle/ptts_oracle.cpp in the absence of reference golden vectors.

It is written the way the upstream *PyTorch* model would be (non-streaming, whole-sequence ops:
causal attention over the full sequence, F.conv1d with left zero padding, F.conv_transpose1d trimmed on
the right, interleaved complex RoPE), i.e. structurally different from the oracle's streaming
per-frame recurrences, but with the same rounding points as ggml's CPU backend (bf16 activations
into bf16 linears, f16 im2col convs, bf16 Mimi KV/q/probs, f16-table GELU) so that agreement is tight.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

torch.set_grad_enabled(False)


def bf(x):
    return x.to(torch.bfloat16).to(torch.float32)


def f16(x):
    return x.to(torch.float16).to(torch.float32)


def gelu_ggml(x):
    y = f16(F.gelu(f16(x), approximate="tanh"))
    y = torch.where(x <= -10.0, torch.zeros_like(x), y)
    return torch.where(x >= 10.0, x, y)


class SecondOpinion:
    def __init__(self, weights: dict, file_bf16: bool = True):
        self.W = {k: torch.from_numpy(np.ascontiguousarray(v)).float() for k, v in weights.items()}
        self.act = bf if file_bf16 else (lambda x: x)

    def lin(self, name, x):
        w = self.W[name + ".weight"]
        b = self.W.get(name + ".bias")
        return F.linear(self.act(x), w, b)

    # ---------------- FlowLM: non-streaming causal transformer over all rows ----------------
    @staticmethod
    def rope_interleaved(x, pos):
        # x [T,H,64]: standard complex rotation of pairs (2i,2i+1); output stays interleaved.
        T, H, D = x.shape
        inv = torch.exp(torch.arange(D // 2, dtype=torch.float32) * (-math.log(10000.0) / (D // 2)))
        ang = pos.to(torch.float32)[:, None] * inv[None, :]
        c, s = torch.cos(ang)[:, None, :], torch.sin(ang)[:, None, :]
        xr, xi = x[..., 0::2], x[..., 1::2]
        out = torch.empty_like(x)
        out[..., 0::2] = xr * c - xi * s
        out[..., 1::2] = xr * s + xi * c
        return out

    def flowlm(self, x):
        """x [T,1024] rows at positions 0..T-1 -> hidden [T,1024] (pre out_norm)."""
        T = x.shape[0]
        pos = torch.arange(T)
        mask = torch.full((T, T), float("-inf")).triu(1)
        for l in range(6):
            p = f"flow_lm.transformer.layers.{l}."
            n = F.layer_norm(x, (1024,), self.W[p + "norm1.weight"], self.W[p + "norm1.bias"], 1e-5)
            qkv = self.lin(p + "self_attn.in_proj", n).view(T, 3, 16, 64)
            q, k, v = self.rope_interleaved(qkv[:, 0], pos), self.rope_interleaved(qkv[:, 1], pos), qkv[:, 2]
            sc = torch.einsum("thd,shd->hts", q, k) / 8.0 + mask[None]
            a = torch.softmax(sc, dim=-1)
            o = torch.einsum("hts,shd->thd", a, v).reshape(T, 1024)
            x = x + self.lin(p + "self_attn.out_proj", o)
            n = F.layer_norm(x, (1024,), self.W[p + "norm2.weight"], self.W[p + "norm2.bias"], 1e-5)
            x = x + self.lin(p + "linear2", gelu_ggml(self.lin(p + "linear1", n)))
        return x

    def t_combined(self):
        out = 0
        for idx, t in ((1, 1.0), (0, 0.0)):
            p = f"flow_lm.flow_net.time_embed.{idx}."
            a = self.W[p + "freqs"] * t
            e = torch.cat([torch.cos(a), torch.sin(a)])
            u = self.lin(p + "mlp.2", F.silu(self.lin(p + "mlp.0", e)))
            u = u / torch.sqrt(u.var(unbiased=True) + 1e-5) * self.W[p + "mlp.3.alpha"]
            out = out + u
        return out / 2

    def head(self, h, noise):
        """h [N,1024] hidden rows, noise [N,32] -> latent [N,32], eos logit+4 [N]."""
        c = F.layer_norm(h, (1024,), self.W["flow_lm.out_norm.weight"], self.W["flow_lm.out_norm.bias"], 1e-5)
        eos = self.lin("flow_lm.out_eos", c)[:, 0] + 4.0
        f = "flow_lm.flow_net."
        x = self.lin(f + "input_proj", noise)
        y = self.t_combined()[None] + self.lin(f + "cond_embed", c)
        sy = F.silu(y)
        for r in range(6):
            p = f + f"res_blocks.{r}."
            shift, scale, gate = self.lin(p + "adaLN_modulation.1", sy).chunk(3, dim=-1)
            hN = F.layer_norm(x, (512,), self.W.get(p + "in_ln.weight"), self.W.get(p + "in_ln.bias"), 1e-6)
            hN = hN * (1 + scale) + shift
            x = x + gate * self.lin(p + "mlp.2", F.silu(self.lin(p + "mlp.0", hN)))
        shift, scale = self.lin(f + "final_layer.adaLN_modulation.1", sy).chunk(2, dim=-1)
        hN = F.layer_norm(x, (512,), self.W.get(f + "final_layer.norm_final.weight"),
                          self.W.get(f + "final_layer.norm_final.bias"), 1e-6)
        v = self.lin(f + "final_layer.linear", hN * (1 + scale) + shift)
        return noise + v, eos

    # ---------------- Mimi: whole-sequence decode of N latents ----------------
    def conv(self, name, x, k):
        """x [C,T] -> causal conv (left zero pad k-1), f16 operands."""
        w = f16(self.W[name + ".conv.weight"])
        b = self.W.get(name + ".conv.bias")
        return F.conv1d(f16(F.pad(x, (k - 1, 0)))[None], w, b)[0]

    def convtr(self, name, x, k, s):
        w = self.W[name + ".convtr.weight"]
        b = self.W.get(name + ".convtr.bias")
        y = F.conv_transpose1d(x[None], w, b, stride=s)[0]
        return y[:, : x.shape[1] * s]

    def mimi_attention_bias(self, offset):
        """[16,250] additive bias per SURVEY.md Appendix D.1 (verbal description, not the pattern code)."""
        bias = torch.zeros(16, 250)
        if offset <= 250:
            for j in range(16):
                for c in range(250):
                    # newest position <= offset+15 living in slot c
                    last = offset + 15
                    p = last - ((last - c) % 250)
                    visible = p >= 0 and p <= offset + j
                    if not visible:
                        bias[j, c] = float("-inf")
        else:
            r = offset % 250
            for j in range(16):
                for d in range(r - 15 + j, r):
                    bias[j, d % 250] = float("-inf")
        return bias

    def mimi(self, latents):
        """latents [N,32] -> pcm [N*1920]."""
        N = latents.shape[0]
        z = latents * self.W["flow_lm.emb_std"] + self.W["flow_lm.emb_mean"]
        e = F.conv1d(f16(z.t())[None], f16(self.W["mimi.quantizer.output_proj.weight"]))[0]       # [512,N]
        up = F.conv_transpose1d(e[None], self.W["mimi.upsample.convtr.convtr.weight"],
                                self.W.get("mimi.upsample.convtr.convtr.bias"), stride=16, groups=512)[0][:, : N * 16]
        x = up.t().contiguous()                                                                      # [N*16,512]
        Tt = N * 16
        pos = torch.arange(Tt)
        for l in range(2):
            p = f"mimi.decoder_transformer.transformer.layers.{l}."
            n = F.layer_norm(x, (512,), self.W[p + "norm1.weight"], self.W[p + "norm1.bias"], 0.0)
            qkv = self.lin(p + "self_attn.in_proj", n).view(Tt, 3, 8, 64)
            q = bf(self.rope_interleaved(qkv[:, 0], pos))
            k = bf(self.rope_interleaved(qkv[:, 1], pos))
            v = bf(qkv[:, 2])
            o = torch.zeros(Tt, 8, 64)
            for f_ in range(N):                      # ring semantics frame by frame (attention only)
                off = 16 * f_
                last = off + 15
                slot_pos = torch.tensor([last - ((last - c) % 250) for c in range(250)])
                valid = slot_pos >= 0
                kk = torch.zeros(250, 8, 64); vv = torch.zeros(250, 8, 64)
                kk[valid] = k[slot_pos[valid]]; vv[valid] = v[slot_pos[valid]]
                sc = torch.einsum("thd,shd->hts", q[off:off + 16], kk) / 8.0 + self.mimi_attention_bias(off)[None]
                a = bf(torch.softmax(sc, dim=-1))
                o[off:off + 16] = torch.einsum("hts,shd->thd", a, vv)
            x = x + self.W[p + "layer_scale_1.scale"] * self.lin(p + "self_attn.out_proj", o.reshape(Tt, 512))
            n = F.layer_norm(x, (512,), self.W[p + "norm2.weight"], self.W[p + "norm2.bias"], 0.0)
            x = x + self.W[p + "layer_scale_2.scale"] * self.lin(p + "linear2", gelu_ggml(self.lin(p + "linear1", n)))
        d = "mimi.decoder.model."
        y = x.t().contiguous()                                                                       # [512, N*16]
        y = F.elu(self.conv(d + "0", y, 7))
        y = self.convtr(d + "2", y, 12, 6)
        y = y + self.conv(d + "3.block.3", F.elu(self.conv(d + "3.block.1", F.elu(y), 3)), 1)
        y = self.convtr(d + "5", F.elu(y), 10, 5)
        y = y + self.conv(d + "6.block.3", F.elu(self.conv(d + "6.block.1", F.elu(y), 3)), 1)
        y = self.convtr(d + "8", F.elu(y), 8, 4)
        y = y + self.conv(d + "9.block.3", F.elu(self.conv(d + "9.block.1", F.elu(y), 3)), 1)
        y = self.conv(d + "11", F.elu(y), 3)
        return y[0]
