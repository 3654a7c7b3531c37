"""clock64 phase trace of CTA (0,0,0) of the pair-schedule ViT attention kernel (both softmax warpgroups), production shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
dbg = torch.zeros(512, dtype=torch.int64, device="cuda")
os.environ["VCB_VIT_TRACE"] = hex(dbg.data_ptr())
os.environ["VCB_VIT_ATTN"] = "pair"
from vcoder_b200 import _lib
lib = _lib.load()
B, S, heads = 24, 577, 16
D = heads * 64
qkv = torch.randn(B * S, 3 * D, device="cuda").half()
o = torch.empty(B * S, D, device="cuda", dtype=torch.float16)
for _ in range(3):
    lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), None)
torch.cuda.synchronize()
d = dbg.cpu()
t0 = int(d[240])
print("cta cycles", int(d[241]) - t0)
names = ["half_start", "S_ready", "S_loaded", "max_done", "Pbuf_free", "rescaled", "exp_P_done", "P_arrived"]
for it in range(2):  # the first two items of CTA 0 (persistent kernel)
    for t in range(2):
        print("item", it, "tile", "AB"[t], " ".join(f"{n:>12s}" for n in names))
        for n in range(10):
            print("   ", n, " ".join(f"{int(d[it * 256 + (t * 10 + n) * 8 + k]) - t0:12d}" for k in range(8)))
        print("    epilogue: O complete", int(d[it * 256 + 200 + t * 2]) - t0, " rows stored", int(d[it * 256 + 201 + t * 2]) - t0)
