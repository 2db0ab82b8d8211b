"""CPU interpreter of a launch plan (test infrastructure).

Executes ``Plan.ops`` with torch CPU ops, using exactly the buffer / channel-offset /
packed-weight semantics the CUDA plan executor implements (csrc/conv_simt.cu
``lp_run_plan``).  It lets the CPU test-suite prove that the *plan* (views, concat
offsets, shuffle strides, BN folding, weight packing) reproduces the oracle, so a GPU
mismatch can only be a kernel bug.  It is never used by the product.
"""
import numpy as np
import torch
import torch.nn.functional as F

from litepi_b200 import _lib as L


def run_plan_cpu(plan, x_u8: np.ndarray):
    """x_u8: [B,S,S,3] RGB uint8.  Returns (buffers list, logits or None)."""
    B = x_u8.shape[0]
    bufs = [None if b["fmt"] == L.FMT_U8 else torch.zeros(B, b["h"], b["w"], b["c"], dtype=torch.float32)
            for b in plan.bufs]
    W = torch.from_numpy(plan.weights())
    logits = None
    for op in plan.ops:
        k, s, cin, cout, cs = op["ksize"], op["stride"], op["cin"], op["cout"], max(op["out_cstride"], 1)
        kind = op["kind"]
        if kind == L.OP_STEM_U8:
            xin = torch.from_numpy(x_u8.astype(np.float32)) / 255.0
            xin = (xin - op["in_mean"]) / op["in_std"]
        else:
            xin = bufs[op["in_buf"]][..., op["in_coff"]:op["in_coff"] + cin]
        if kind in (L.OP_STEM_U8, L.OP_CONV):
            w = W[op["w_off"]:op["w_off"] + k * k * cin * cout].reshape(k, k, cin, cout).permute(3, 2, 0, 1)
            b = W[op["b_off"]:op["b_off"] + cout]
            y = F.conv2d(xin.permute(0, 3, 1, 2), w, b, stride=s, padding=k // 2).permute(0, 2, 3, 1)
            if op["act"] == L.ACT_SILU:
                y = y * torch.sigmoid(y)
            elif op["act"] == L.ACT_RELU:
                y = torch.relu(y)
            if op["res_buf"] >= 0:
                y = y + bufs[op["res_buf"]][..., op["res_coff"]:op["res_coff"] + cout]
        elif kind == L.OP_DWCONV3:
            w = W[op["w_off"]:op["w_off"] + 9 * cout].reshape(3, 3, cout).permute(2, 0, 1).unsqueeze(1)
            b = W[op["b_off"]:op["b_off"] + cout]
            y = F.conv2d(xin.permute(0, 3, 1, 2), w, b, stride=s, padding=1, groups=cout).permute(0, 2, 3, 1)
        elif kind == L.OP_MAXPOOL:
            y = F.max_pool2d(xin.permute(0, 3, 1, 2), k, s, k // 2).permute(0, 2, 3, 1)
        elif kind == L.OP_UPSAMPLE2:
            y = xin.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2)
        elif kind == L.OP_COPY:
            y = xin
        elif kind == L.OP_MEAN_FC:
            w = W[op["w_off"]:op["w_off"] + cin * cout].reshape(cin, cout)
            b = W[op["b_off"]:op["b_off"] + cout]
            logits = xin.mean(dim=(1, 2)) @ w + b
            continue
        else:
            raise ValueError(kind)
        ob = bufs[op["out_buf"]]
        if plan.bufs[op["out_buf"]]["w"] == 1 and plan.bufs[op["out_buf"]]["h"] > 1:      # Detect head rows
            rows = y.shape[1] * y.shape[2]
            ob[:, op["row_off"]:op["row_off"] + rows, 0, op["out_coff"]:op["out_coff"] + cout] = y.reshape(B, rows, cout)
        elif op.get("out_seg_len", 0) > 0:                                 # segmented (shuffled) destination
            n = op["cout_real"]
            l = op["out_coff"] + cs * np.arange(n)
            phys = (l // op["out_seg_len"]) * op["out_seg_pad"] + l % op["out_seg_len"]
            ob[..., torch.from_numpy(phys)] = y[..., :n]
        else:
            ob[..., op["out_coff"]:op["out_coff"] + cs * cout:cs] = y
    return bufs, logits


def detect_tail_cpu(head: torch.Tensor, in_size: int = 640):
    """head [B,A,HC] -> out0 [B,5,A]; Detect tail of model.ncnn.param:184-208 (nc = 1)."""
    B, A, _ = head.shape
    d = torch.softmax(head[..., :64].reshape(B, A, 4, 16), dim=-1) @ torch.arange(16, dtype=torch.float32)
    ax, ay, st = [], [], []
    for s in (8, 16, 32):
        n = in_size // s
        ys, xs = torch.meshgrid(torch.arange(n) + 0.5, torch.arange(n) + 0.5, indexing="ij")
        ax.append(xs.reshape(-1)); ay.append(ys.reshape(-1)); st.append(torch.full((n * n,), float(s)))
    ax, ay, st = torch.cat(ax), torch.cat(ay), torch.cat(st)
    x1, y1, x2, y2 = ax - d[..., 0], ay - d[..., 1], ax + d[..., 2], ay + d[..., 3]
    out = torch.stack([(x1 + x2) / 2 * st, (y1 + y2) / 2 * st, (x2 - x1) * st, (y2 - y1) * st,
                       torch.sigmoid(head[..., 64])], dim=1)
    return out
