"""Host-side operand layouts of the generator's conv stack (audiogan_b200/plan.py) checked on the CPU: the packed filter of a
channel-prefix conv in both column orders of include/audiogan_b200.h (a_layout 1: (channel group of 8, tap padded to 8, channel);
a_layout 2: (tap, channel group of 64, channel)) is a pure re-indexing of the canonical weight-normed filter
(audiogan.py:266-283), and the gradient map back (q2c) is its inverse.  No kernel runs here: the gather is emulated with
torch indexing on the plan's own index tables."""
import torch

import audiogan_b200 as ag
from audiogan_b200 import plan as P


def _packed(pl, w, name):
    idx = pl._pack_idx[name]
    return torch.where(idx >= 0, w[idx.clamp(min=0)], torch.zeros(()))


def test_prefix_conv_filter_layouts_are_reindexings():
    g = ag.Generator(embed_size=100)
    pl = P.build_generator_plan(g, torch.device("cpu"))
    assert pl.alay == [1, 2, 2, 2] and pl.cinp == [16, 32, 64, 96]          # default net: 16-channel slots, 128-byte boxes from 32 up
    w = torch.randn(pl.canon.size)
    for li, (k, s, hid, out) in enumerate(g._struct):
        cp = pl.cinp[li]
        wc = _packed(pl, w, "c%d.w" % li).view(hid, k, cp)                  # (h, tap, padded channel)
        wq = _packed(pl, w, "c%d.wq" % li)
        assert wq.shape == (hid, pl.Kq[li])
        if pl.alay[li] == 2:
            G = (cp + 63) // 64
            assert pl.Kq[li] == k * G * 64
            q = wq.view(hid, k, G * 64)
            assert torch.equal(q[:, :, :cp], wc)
            assert G * 64 == cp or float(q[:, :, cp:].abs().max()) == 0     # columns of channels past the prefix hold zeros
        else:
            G, KT = cp // 8, (k + 7) // 8
            assert pl.Kq[li] == G * KT * 64
            q = wq.view(hid, G, KT * 8, 8)
            assert torch.equal(q[:, :, :k].permute(0, 2, 1, 3).reshape(hid, k, cp), wc)
            assert KT * 8 == k or float(q[:, :, k:].abs().max()) == 0       # padded taps hold zeros
        # canonical channels only: the pad channels of a slot are structural zeros in both layouts
        real = int((wc[0, 0] != 0).sum())
        assert real == 1 + sum(o for (_, _, _, o) in g._struct[:li])


def test_prefix_conv_gradient_map_inverts_the_layout():
    g = ag.Generator(embed_size=100)
    pl = P.build_generator_plan(g, torch.device("cpu"))
    for li, (k, s, hid, out) in enumerate(g._struct):
        cp, Kq = pl.cinp[li], pl.Kq[li]
        gq = torch.randn(hid, Kq + 1)                                       # what the weight-gradient GEMM leaves (+ bias column)
        got = gq.reshape(-1)[pl.q2c[li].long()]                             # K.gather(c.w grad region, c.wq grad region, q2c)
        assert got.shape == (hid, k * cp + 1)
        if pl.alay[li] == 2:
            G = (cp + 63) // 64
            ref = gq[:, :Kq].view(hid, k, G * 64)[:, :, :cp].reshape(hid, k * cp)
        else:
            G, KT = cp // 8, (k + 7) // 8
            ref = gq[:, :Kq].view(hid, G, KT * 8, 8)[:, :, :k].permute(0, 2, 1, 3).reshape(hid, k * cp)
        assert torch.equal(got[:, :-1], ref) and torch.equal(got[:, -1], gq[:, Kq])
