"""Torch-CPU restatement of the reference stacked-hourglass CNN (oracle; see oracle/__init__.py).

Follows src/mvlm/prediction/paulsenpredictor.py:
  ResidualBlock.forward   :267-273
  HourGlassModule.forward :301-361
  MVLMModel.forward       :404-432  (eval mode: BN running stats, dropout = identity;
                                     only outputs[-1] = conv11 branch is consumed, :204-205)
It is written functionally over a state_dict (keys as in mvlm_b200/weights.py) and in
the same producer/consumer order as the CUDA plan so that one `q` hook can emulate
the bf16 storage points of the CUDA path:
  q = identity            -> fp32 oracle (pinned against the reference in tests/golden)
  q = bf16 round-trip     -> "same rounding points" oracle for tight kernel checks
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


def _ident(t):
    return t


def bf16_round(t):
    return t.to(torch.bfloat16).float()


class HourglassOracle:
    def __init__(self, sd: dict, emulate_bf16: bool = False):
        self.sd = {k: v.detach().float() if v.is_floating_point() else v for k, v in sd.items()}
        self.q = bf16_round if emulate_bf16 else _ident
        self.emulate = emulate_bf16

    # -- helpers ------------------------------------------------------------
    def bn(self, prefix):
        sd = self.sd
        scale = sd[f"{prefix}.weight"] / torch.sqrt(sd[f"{prefix}.running_var"] + BN_EPS)
        shift = sd[f"{prefix}.bias"] - sd[f"{prefix}.running_mean"] * scale
        return scale.view(1, -1, 1, 1), shift.view(1, -1, 1, 1)

    def act(self, x, prefix):
        s, t = self.bn(prefix)
        return self.q(torch.relu(x * s + t))

    def conv(self, x, name, quant_w=True):
        w = self.sd[f"{name}.weight"]
        if quant_w:
            w = self.q(w)
        b = self.sd.get(f"{name}.bias")
        return F.conv2d(x, w, b, padding=w.shape[-1] // 2)

    def rb(self, x_raw, a_in, p, post=None, up_low=None):
        """ResidualBlock (:267-273).  x_raw: stored block input, a_in = relu(bn1(x)) stored.
        up_low: half-resolution tensor added nearest-x2 up-sampled (:334-359), like the CUDA plan's epilogue
        before the block output is rounded.  Returns (y stored, relu(post_bn(y)) stored or None)."""
        sd = self.sd
        if f"{p}.resample.2.weight" in sd:
            ar = self.act(x_raw, f"{p}.resample.0")
            skip = self.q(F.conv2d(ar, self.q(sd[f"{p}.resample.2.weight"])))
        else:
            skip = x_raw
        o1 = self.conv(a_in, f"{p}.conv1")
        a1 = self.act(o1, f"{p}.bn2")
        o2 = self.conv(a1, f"{p}.conv2")
        a2 = self.act(o2, f"{p}.bn3")
        o3 = self.conv(a2, f"{p}.conv3")
        y = torch.cat((o1, o2, o3), 1) + skip
        if up_low is not None:
            y = y + F.interpolate(up_low, scale_factor=2, mode="nearest")
        post_act = self.act(y, post) if post is not None else None  # from the un-rounded sum
        return self.q(y), post_act

    def hourglass(self, x, a_x, p):
        """HourGlassModule.forward (:301-361). x stored raw input, a_x = relu(rb1.bn1(x)).
        Same schedule as the CUDA plan: the skip-branch blocks (rb1,3,5,7,9) run on the way up with
        `interpolate(low) + skip` folded into their output (one rounding instead of two)."""
        cur = x
        low_blocks = [2, 4, 6, 8, 10]
        skip_blocks = [3, 5, 7, 9]
        lows, a_lows = [], []
        for lvl, lb in enumerate(low_blocks):
            pooled = F.max_pool2d(cur, 2)
            a = self.act(pooled, f"{p}.rb{lb}.bn1")
            nxt = skip_blocks[lvl] if lvl < 4 else 11
            low, a_low = self.rb(pooled, a, f"{p}.rb{lb}", post=f"{p}.rb{nxt}.bn1")
            lows.append(low)
            a_lows.append(a_low)
            cur = low
        # bottleneck (:332-334)
        low2, a2 = self.rb(cur, a_lows[4], f"{p}.rb11", post=f"{p}.rb12.bn1")
        low3, _ = self.rb(low2, a2, f"{p}.rb12")
        cur = low3
        # up path (:335-359)
        for lvl, (b1, b2) in enumerate([(13, 14), (15, 16), (17, 18), (19, 20)]):
            s, a = self.rb(lows[3 - lvl], a_lows[3 - lvl], f"{p}.rb{skip_blocks[3 - lvl]}", post=f"{p}.rb{b1}.bn1", up_low=cur)
            l1, a1 = self.rb(s, a, f"{p}.rb{b1}", post=f"{p}.rb{b2}.bn1")
            cur, _ = self.rb(l1, a1, f"{p}.rb{b2}")
        add5, _ = self.rb(x, a_x, f"{p}.rb1", up_low=cur)
        return add5

    # -- full forward ---------------------------------------------------------
    def forward(self, img_nchw: torch.Tensor, return_intermediates: bool = False):
        """img (B,Cin,H,W) fp32 in [0,1] -> heatmaps (B,L,H,W) fp32 (= outputs[-1] of the reference)."""
        sd = self.sd
        with torch.no_grad():
            inter = {}
            # stem (:405-407): the CUDA path keeps the image to ~2^-17 (hi/lo bf16 split) and rounds only the
            # conv1 weights to bf16, like every other layer; x0 itself is never stored
            x0 = torch.relu(self._bn_apply(F.conv2d(img_nchw.float(), self.q(sd["conv1.weight"]), sd["conv1.bias"], padding=1), "bn1"))
            a = self.act(x0, "conv2.bn1")
            # conv2 block has a resample skip that consumes x0 through its own BN
            y2, _ = self._rb_from_fp32(x0, a, "conv2")
            x1 = F.max_pool2d(y2, 2)
            inter["x1"] = x1
            a = self.act(x1, "conv3.bn1")
            y3, a4 = self.rb(x1, a, "conv3", post="conv4.bn1")
            inter["y3"] = y3
            r3, a_h1 = self.rb(y3, a4, "conv4", post="hg1.rb1.bn1")
            inter["r3"] = r3
            h1 = self.hourglass(r3, a_h1, "hg1")
            inter["hg1"] = h1
            ll1 = self.act(self.conv(h1, "conv5"), "bn2")
            x6 = self.q(self.conv(ll1, "conv6"))
            s = self.conv(x6, "conv7") + r3 + ll1
            a_h2 = self.act(s, "hg2.rb1.bn1")
            s = self.q(s)
            inter["sum_temp"] = s
            h2 = self.hourglass(s, a_h2, "hg2")
            ll2 = self.act(self.conv(h2, "conv9"), "bn3")
            x10 = self.q(self.conv(ll2, "conv10"))
            inter["x10"] = x10
            if self.emulate:
                hm = self._conv11_phases(x10)
            else:
                up = F.interpolate(x10, scale_factor=2, mode="nearest")
                hm = self.conv(up, "conv11")
            if return_intermediates:
                return hm, inter
            return hm

    def _bn_apply(self, x, prefix):
        s, t = self.bn(prefix)
        return x * s + t

    def _rb_from_fp32(self, x0, a_in, p):
        """conv2 block: its input x0 is never stored (stem output stays fp32 inside the stem kernel)."""
        sd = self.sd
        ar = self.act(x0, f"{p}.resample.0")
        skip = self.q(F.conv2d(ar, self.q(sd[f"{p}.resample.2.weight"])))
        o1 = self.conv(a_in, f"{p}.conv1")
        a1 = self.act(o1, f"{p}.bn2")
        # in the CUDA plan each slice is accumulated in place into the stored skip tensor
        y1 = self.q(o1 + skip[:, : o1.shape[1]])
        o2 = self.conv(a1, f"{p}.conv2")
        a2 = self.act(o2, f"{p}.bn3")
        c1 = o1.shape[1]
        c2 = c1 + o2.shape[1]
        y2 = self.q(o2 + skip[:, c1:c2])
        o3 = self.conv(a2, f"{p}.conv3")
        y3 = self.q(o3 + skip[:, c2:])
        return torch.cat((y1, y2, y3), 1), None

    def _conv11_phases(self, x10):
        """conv3x3(nearest_up2(x)) as four 2x2 phase convolutions at low resolution, the taps that
        read the same low-res pixel summed in fp32 and THEN rounded to bf16 (what the CUDA plan does;
        mathematically identical to :428-429, differs only in where the weight rounding happens)."""
        w = self.sd["conv11.weight"]
        b = self.sd["conv11.bias"]
        n, _, h, wd = x10.shape
        out = torch.empty((n, w.shape[0], 2 * h, 2 * wd))
        for a in range(2):
            wr = torch.stack([w[:, :, 0], w[:, :, 1] + w[:, :, 2]], 2) if a == 0 else \
                torch.stack([w[:, :, 0] + w[:, :, 1], w[:, :, 2]], 2)
            for c in range(2):
                wc = torch.stack([wr[..., 0], wr[..., 1] + wr[..., 2]], 3) if c == 0 else \
                    torch.stack([wr[..., 0] + wr[..., 1], wr[..., 2]], 3)
                # tap (ky,kx) reads input (y + a - 1 + ky, x + c - 1 + kx)
                xp = F.pad(x10, (1 - c, c, 1 - a, a))
                out[:, :, a::2, c::2] = F.conv2d(xp, self.q(wc), b)
        return out
