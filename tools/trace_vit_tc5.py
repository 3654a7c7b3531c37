"""Print the clock64 phase trace of CTA (0,0,0) of the tcgen05 ViT attention kernel."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
dbg = torch.zeros(64, dtype=torch.int64, device="cuda")
os.environ["VCB_VIT_TRACE"] = hex(dbg.data_ptr())
from vcoder_b200 import _lib
lib = _lib.load()
B, S, heads = 24, 577, 16
D = heads * 64
qkv = torch.randn(B * S, 3 * D, device="cuda").half()
o = torch.empty(B * S, D, device="cuda", dtype=torch.float16)
for _ in range(3):
    lib.vcb_op_attention_vit_f16(_lib.ptr(qkv), B, S, heads, 0.125, _lib.ptr(o), None)
torch.cuda.synchronize()
t = dbg.cpu().view(8, 8)[:5]
t0 = int(t[0, 0])
names = ["blk_start", "S_ready", "S_loaded", "PVprev_done", "after_xchg", "P_arrived", "mma_saw_P", "mma_issued_PV"]
print("   " + " ".join(f"{n:>13s}" for n in names))
for j in range(5):
    print(j, " ".join(f"{int(t[j, k]) - t0:13d}" for k in range(8)))
