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


def _act(y, act):
    if act == L.ACT_SILU:
        return y * torch.sigmoid(y)
    if act == L.ACT_RELU:
        return torch.relu(y)
    if act == L.ACT_RELU6:
        return torch.clamp(y, 0.0, 6.0)
    if act == L.ACT_SIGMOID:
        return torch.sigmoid(y)
    return y


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
            res_first = bool(op.get("flags", 0) & L.OPF_RES_BEFORE_ACT)
            if op["res_buf"] >= 0 and res_first:
                y = y + bufs[op["res_buf"]][..., op["res_coff"]:op["res_coff"] + cout]
            y = _act(y, op["act"])
            if op["res_buf"] >= 0 and not res_first:
                y = y + bufs[op["res_buf"]][..., op["res_coff"]:op["res_coff"] + cout]
        elif kind == L.OP_DWCONV3:
            w = W[op["w_off"]:op["w_off"] + k * k * cout].reshape(k, k, cout).permute(2, 0, 1).unsqueeze(1)
            b = W[op["b_off"]:op["b_off"] + cout]
            y = _act(F.conv2d(xin.permute(0, 3, 1, 2), w, b, stride=s, padding=k // 2, groups=cout).permute(0, 2, 3, 1), op["act"])
        elif kind == L.OP_GLOBAL_MEAN:
            y = xin.mean(dim=(1, 2), keepdim=True)
        elif kind == L.OP_SCALE:
            y = xin * bufs[op["res_buf"]][..., op["res_coff"]:op["res_coff"] + cout]
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


def run_fused_cpu(prog, x_u8, grid: int = 2, mean=0.18, std=0.34):
    """CPU execution of the fused-classifier program (plan.FusedProgram) with the SAME shared-memory maps (flat float
    arrays, NaN-initialised), so buffer overlays and offsets are validated, not just the math.  Mirrors the kernel's
    schedule: ``grid`` CTAs, CTA b owns ROIs b, b + grid, ...; per ROI front then middle (parked in "global"
    memory), then the tail over up to tail_group of the CTA's ROIs stacked as rows.  x_u8 [R,64,64,3]."""
    steps, W = prog.steps, np.asarray(prog.weights, np.float32)
    R, GT = x_u8.shape[0], prog.tail_group
    norm = ((np.arange(256, dtype=np.float32) / np.float32(255) - np.float32(mean)) / np.float32(std)).astype(np.float32)
    logits = np.full((R, int(steps[-1][9])), np.nan, np.float32)

    def run_steps(sm, rng, rois, roi_ids, park):
        for si in rng:
            (op, src, dst, src_C, src_off, dst_C, dst_off, dst_cs, cin, cout, H, Wd, stride, relu, w_off, b_off,
             roi_stride, _) = [int(v) for v in steps[si]]
            if op in (2, 4, 6, 7):
                Ho, Wo = H, Wd
            else:
                Ho, Wo = (H + 2 - 3) // stride + 1, (Wd + 2 - 3) // stride + 1
            if op == 0:        # conv1 from the u8 crop
                img = torch.from_numpy(norm[x_u8[roi_ids[0]]]).permute(2, 0, 1)[None]
                cp = (cout + 3) // 4 * 4
                w = torch.from_numpy(W[w_off:w_off + 27 * cp].reshape(3, 3, 3, cp)[..., :cout]).permute(3, 2, 0, 1)
                y = torch.relu(F.conv2d(img, w, torch.from_numpy(W[b_off:b_off + cout]), stride=2, padding=1))
                sm[dst:dst + Ho * Wo * dst_C] = y[0].permute(1, 2, 0).reshape(-1).numpy()
                continue
            if op == 7:        # load the parked tensors of the stacked ROIs
                for g, rid in enumerate(roi_ids):
                    sm[dst + g * roi_stride:dst + (g + 1) * roi_stride] = park[rid]
                continue
            src_t = torch.from_numpy(sm[src:src + rois * H * Wd * src_C].reshape(rois, H, Wd, src_C).copy())
            if op == 6:        # park the middle's result
                assert not torch.isnan(src_t).any()
                park[roi_ids[0]] = src_t.reshape(-1).numpy().copy()
                continue
            if op == 1:
                y = F.max_pool2d(src_t[..., src_off:src_off + cout].permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1)
            elif op == 2:
                cp = (cout + 3) // 4 * 4
                w = torch.from_numpy(W[w_off:w_off + cin * cp].reshape(cin, cp)[:, :cout])
                y = src_t[..., src_off:src_off + cin] @ w + torch.from_numpy(W[b_off:b_off + cout])
                if relu:
                    y = torch.relu(y)
            elif op == 3:
                w = torch.from_numpy(W[w_off:w_off + 9 * cout].reshape(3, 3, cout)).permute(2, 0, 1).unsqueeze(1)
                y = F.conv2d(src_t[..., src_off:src_off + cout].permute(0, 3, 1, 2), w, torch.from_numpy(W[b_off:b_off + cout]),
                             stride=stride, padding=1, groups=cout).permute(0, 2, 3, 1)
            elif op == 4:
                y = src_t[..., src_off:src_off + cout]
            elif op == 5:
                m = src_t[..., :cin].mean(dim=(1, 2))
                w = torch.from_numpy(W[w_off:w_off + cin * cout].reshape(cin, cout))
                out = (m @ w + torch.from_numpy(W[b_off:b_off + cout])).numpy()
                for g, rid in enumerate(roi_ids):
                    logits[rid] = out[g]
                continue
            assert not torch.isnan(y).any(), f"step {si} reads uninitialised shared memory"
            d = sm[dst:dst + rois * Ho * Wo * dst_C].reshape(rois, Ho, Wo, dst_C)
            d[..., dst_off:dst_off + dst_cs * cout:dst_cs] = y.numpy()

    nf, nm = prog.n_front, prog.n_mid
    for b in range(grid):
        mine = list(range(b, R, grid))
        for k0 in range(0, len(mine), GT):
            chunk = mine[k0:k0 + GT]
            park = {}
            for rid in chunk:
                sm = np.full(prog.smem_bytes // 4, np.nan, np.float32)
                run_steps(sm, range(0, nf), 1, [rid], park)
                run_steps(sm, range(nf, nf + nm), 1, [rid], park)
            sm = np.full(prog.tail_bytes // 4, np.nan, np.float32)
            run_steps(sm, range(nf + nm, len(steps)), len(chunk), chunk, park)
    assert not np.isnan(logits).any()
    return logits
