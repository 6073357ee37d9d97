import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import feature_level_style_transfer_for_tsc_b200 as T
from feature_level_style_transfer_for_tsc_b200 import ops
L = T._lib
B, C, Ln = 128, 144, 128
a, s = torch.randn(B, C, Ln, device="cuda"), torch.randn(B, C, Ln, device="cuda")
tl = torch.zeros(64, device="cuda", dtype=torch.int64)
ops.gram_loss_fwd(L.ENGINE_TCGEN05, a, s); torch.cuda.synchronize()
L.load().tsc_debug_set_timeline(tl.data_ptr())
ops.gram_loss_fwd(L.ENGINE_TCGEN05, a, s); torch.cuda.synchronize()
L.load().tsc_debug_set_timeline(None)
t = tl.cpu().tolist()
print("acc ready", t[5] - t[0], "epilogue done", t[6] - t[0])
print("mma chunk ready:", [v - t[0] for v in t[8:24] if v])
print("producer chunk filled:", [v - t[0] for v in t[24:40] if v])
