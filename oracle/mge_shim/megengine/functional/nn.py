"""F.nn.roi_align (ASSUMED-6: aligned -> offset 0.5, zero padding outside the map, a + (b-a)*t lerp, sum / S^2)."""
import numpy as np

from ..tensor import Tensor

f32 = np.float32


def roi_align(inp, rois, output_shape, mode="average", spatial_scale=1.0, sample_points=2, aligned=True):
    assert mode == "average"
    feat = inp._a.astype(f32)
    r = rois._a.astype(f32)
    if isinstance(output_shape, int):
        output_shape = (output_shape, output_shape)
    if isinstance(sample_points, int):
        sample_points = (sample_points, sample_points)
    PH, PW = output_shape
    SH, SW = sample_points
    _, C, H, W = feat.shape
    scale = f32(spatial_scale)
    off = f32(0.5) if aligned else f32(0.0)
    out = np.zeros((r.shape[0], C, PH, PW), dtype=f32)
    for k in range(r.shape[0]):
        fm = feat[int(r[k, 0])]
        sw_, sh_ = f32(r[k, 1] * scale - off), f32(r[k, 2] * scale - off)
        ew_, eh_ = f32(r[k, 3] * scale - off), f32(r[k, 4] * scale - off)
        rw = max(f32(ew_ - sw_), f32(0))
        rh = max(f32(eh_ - sh_), f32(0))
        bh, bw = f32(rh / f32(PH)), f32(rw / f32(PW))
        for ph in range(PH):
            for pw in range(PW):
                acc = np.zeros(C, dtype=f32)
                for iy in range(SH):
                    for ix in range(SW):
                        hc = f32(sh_ + f32(bh * f32(f32(ph) + f32(f32(iy + 0.5) / f32(SH)))))
                        wc = f32(sw_ + f32(bw * f32(f32(pw) + f32(f32(ix + 0.5) / f32(SW)))))
                        h0, w0 = int(np.floor(hc)), int(np.floor(wc))
                        h1, w1 = h0 + 1, w0 + 1

                        def tap(y, x):
                            return fm[:, y, x] if (0 <= y < H and 0 <= x < W) else np.zeros(C, dtype=f32)

                        lw, lh = f32(wc - f32(w0)), f32(hc - f32(h0))
                        tl, tr, bl, br = tap(h0, w0), tap(h0, w1), tap(h1, w0), tap(h1, w1)
                        top = (tl + ((tr - tl).astype(f32) * lw).astype(f32)).astype(f32)
                        bot = (bl + ((br - bl).astype(f32) * lw).astype(f32)).astype(f32)
                        acc = (acc + (top + ((bot - top).astype(f32) * lh).astype(f32)).astype(f32)).astype(f32)
                out[k, :, ph, pw] = (acc / f32(SH * SW)).astype(f32)
    return Tensor(out)


def roi_pooling(inp, rois, output_shape, mode="max", scale=1.0):
    raise NotImplementedError("max roi_pooling is outside the hot path")
