"""A/B of the FFN-conv GEMM tile schedule (NS2_GEMM_FLAG_NARROW_LAST) at the cfg2 shape, interleaved launches."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from naturalspeech2_pytorch_b200 import ops
B, N, Di, Dp = 32, 1024, 1365, 1408
bf = torch.bfloat16
g = (torch.randn(B, N, Dp, device="cuda") * 0.5).to(bf)
wc = (torch.randn(Dp, 3 * Dp, device="cuda") * 0.02).to(bf)
bc = torch.randn(Dp, device="cuda")
outs = [torch.empty(B, N, Dp, device="cuda", dtype=bf) for _ in range(2)]
def run(flag, out):
    ops.gemm(g, wc, out, n=Dp, epilogue=ops.EPI_BF16, bias=bc, segs=ops.conv3_segs(Dp), flags=flag)
for f in (0, 8):
    run(f, outs[f // 8])
torch.cuda.synchronize()
print("identical:", torch.equal(outs[0], outs[1]))
for rnd in range(3):
    for f in (0, 8):
        for _ in range(5):
            run(f, outs[0])
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(40):
            run(f, outs[0])
        b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / 40
        print(f"round {rnd} flags={f}: {ms:.4f} ms  {2.0 * B * N * Di * 3 * Di / ms / 1e9:.0f} TFLOP/s")
