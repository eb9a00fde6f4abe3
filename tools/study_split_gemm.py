"""Design study for the round-2 tensor-core loss path (CPU, no kernels): forward error of a 512-deep GEMM when the
operands are rounded to TF32, or split into bf16 hi + lo parts with three products (hi*hi + hi*lo + lo*hi, fp32
accumulation) — the operand formats `tcgen05.mma` offers — against fp64.  The measured loss errors of the TF32 build
(profiles/r01_loss_gemm_variants.log: up to 1.3e-3) calibrate what the GEMM-level numbers mean for the loss fixtures.

    python tests/study_split_gemm.py
"""
import torch


def tf32(x):
    """round-to-nearest onto TF32's 10 explicit mantissa bits"""
    i = x.contiguous().view(torch.int32)
    i = (i + 0x00000FFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


def split_bf16(x):
    hi = x.bfloat16().float()
    return hi, (x - hi).bfloat16().float()


def main():
    torch.manual_seed(0)
    rows, K, N = 4096, 512, 512
    cases = {
        "tanh activations x N(0, 1/sqrt(K)) weights": (torch.tanh(torch.randn(rows, K)), torch.randn(K, N) / K ** 0.5),
        "tangent rows (heavy tails) x weights": (torch.randn(rows, K) * torch.randn(rows, 1).exp(), torch.randn(K, N) / K ** 0.5),
    }
    print(f"{'case':48s} {'fp32':>10s} {'tf32':>10s} {'3xtf32':>10s} {'bf16x3':>10s} {'bf16':>10s}   (relative Frobenius error vs fp64)")
    for name, (a, w) in cases.items():
        ref = a.double() @ w.double()
        nrm = ref.norm()
        err = lambda c: float((c.double() - ref).norm() / nrm)
        ah, al = split_bf16(a)
        wh, wl = split_bf16(w)
        at, wt = tf32(a), tf32(w)
        atl, wtl = tf32(a - at), tf32(w - wt)
        res = {
            "fp32": a @ w,
            "tf32": at @ wt,
            "3xtf32": atl @ wt + at @ wtl + at @ wt,
            "bf16x3": al @ wh + ah @ wl + ah @ wh,
            "bf16": ah @ wh,
        }
        print(f"{name:48s} " + " ".join(f"{err(res[k]):10.2e}" for k in ("fp32", "tf32", "3xtf32", "bf16x3", "bf16")))


if __name__ == "__main__":
    main()
