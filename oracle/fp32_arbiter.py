"""Third, independent evaluation of the INTEGER-deciding chain (depth bits, radius, tile rect) for
a handful of Gaussians, in numpy float32 scalar arithmetic (IEEE, one rounding per operation, no
FMA) in exactly the operation order of oracle/splat_ref.py::preprocess and csrc/project.cu.
TEST INFRASTRUCTURE ONLY: used to arbitrate if the torch oracle and the CUDA kernel ever disagree
on an integer (torch CPU kernels and CUDA intrinsics are both IEEE, so they should not)."""
import numpy as np

f32 = np.float32


def radius_rect(mean, scale, quat, V, F, tanfovx, tanfovy, H, W, scale_modifier=1.0):
    """-> dict(depth_bits, radius, rect=(minx,miny,maxx,maxy), visible).  All inputs array-likes."""
    x, y, z = (f32(v) for v in mean)
    V = np.asarray(V, np.float32).reshape(16)
    F = np.asarray(F, np.float32).reshape(16)
    tx = ((V[0] * x + V[4] * y) + V[8] * z) + V[12]
    ty = ((V[1] * x + V[5] * y) + V[9] * z) + V[13]
    tz = ((V[2] * x + V[6] * y) + V[10] * z) + V[14]
    out = dict(depth_bits=int(np.array([tz], np.float32).view(np.uint32)[0]), radius=0, rect=(0, 0, 0, 0), visible=False)
    if not tz > f32(0.2):
        return out
    hx = ((F[0] * x + F[4] * y) + F[8] * z) + F[12]
    hy = ((F[1] * x + F[5] * y) + F[9] * z) + F[13]
    hw = ((F[3] * x + F[7] * y) + F[11] * z) + F[15]
    pw = f32(1.0) / (hw + f32(1e-7))
    ndcx, ndcy = hx * pw, hy * pw
    Wf, Hf = f32(W), f32(H)
    px = ((ndcx + f32(1.0)) * Wf - f32(1.0)) * f32(0.5)
    py = ((ndcy + f32(1.0)) * Hf - f32(1.0)) * f32(0.5)
    mod = f32(scale_modifier)
    s = [mod * f32(v) for v in scale]
    r, qx, qy, qz = (f32(v) for v in quat)
    two, one = f32(2.0), f32(1.0)
    R = [[one - two * (qy * qy + qz * qz), two * (qx * qy - r * qz), two * (qx * qz + r * qy)],
         [two * (qx * qy + r * qz), one - two * (qx * qx + qz * qz), two * (qy * qz - r * qx)],
         [two * (qx * qz - r * qy), two * (qy * qz + r * qx), one - two * (qx * qx + qy * qy)]]
    L = [[R[i][j] * s[j] for j in range(3)] for i in range(3)]
    dot = lambda i, j: (L[i][0] * L[j][0] + L[i][1] * L[j][1]) + L[i][2] * L[j][2]
    S = [[dot(i, j) for j in range(3)] for i in range(3)]
    tfx, tfy = f32(tanfovx), f32(tanfovy)
    limx, limy = f32(1.3) * tfx, f32(1.3) * tfy
    fx, fy = Wf / (two * tfx), Hf / (two * tfy)
    txtz, tytz = tx / tz, ty / tz
    cx = min(limx, max(-limx, txtz)) * tz
    cy = min(limy, max(-limy, tytz)) * tz
    tz2 = tz * tz
    J00, J02 = fx / tz, -((fx * cx) / tz2)
    J11, J12 = fy / tz, -((fy * cy) / tz2)
    Wr = [[V[4 * k + i] for k in range(3)] for i in range(3)]
    M0 = [J00 * Wr[0][k] + J02 * Wr[2][k] for k in range(3)]
    M1 = [J11 * Wr[1][k] + J12 * Wr[2][k] for k in range(3)]
    N0 = [(M0[0] * S[0][j] + M0[1] * S[1][j]) + M0[2] * S[2][j] for j in range(3)]
    N1 = [(M1[0] * S[0][j] + M1[1] * S[1][j]) + M1[2] * S[2][j] for j in range(3)]
    a = ((N0[0] * M0[0] + N0[1] * M0[1]) + N0[2] * M0[2]) + f32(0.3)
    b = (N0[0] * M1[0] + N0[1] * M1[1]) + N0[2] * M1[2]
    c = ((N1[0] * M1[0] + N1[1] * M1[1]) + N1[2] * M1[2]) + f32(0.3)
    det = a * c - b * b
    if det == 0:
        return out
    mid = f32(0.5) * (a + c)
    sq = np.sqrt(max(mid * mid - det, f32(0.1)), dtype=np.float32)
    lam = max(mid + sq, mid - sq)
    rad_f = np.ceil(f32(3.0) * np.sqrt(lam, dtype=np.float32))
    rad_f = f32(min(max(float(rad_f) if rad_f == rad_f else 0.0, 0.0), 1.0e9))
    gx, gy = (W + 15) // 16, (H + 15) // 16

    def tile(v, g):
        t = f32(v) * f32(0.0625)
        t = f32(0.0) if t != t else t
        t = min(max(t, f32(-1.0)), f32(g + 1))
        return min(g, max(0, int(np.trunc(t))))

    rect = (tile(px - rad_f, gx), tile(py - rad_f, gy), tile((px + rad_f) + f32(15.0), gx),
            tile((py + rad_f) + f32(15.0), gy))
    touched = (rect[2] - rect[0]) * (rect[3] - rect[1])
    if touched > 0:
        out.update(radius=int(rad_f), rect=rect, visible=True)
    out["cov2d"] = (float(a), float(b), float(c))
    return out
