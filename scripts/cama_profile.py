import sys
from pathlib import Path
import torch, torch.nn as nn
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import motionrag_b200 as m
b = int(sys.argv[1]) if len(sys.argv) > 1 else 1
torch.manual_seed(0)
layer = nn.TransformerEncoderLayer(1024, 16, 4096, 0.0, "gelu", batch_first=True, norm_first=False, bias=True)
enc = nn.TransformerEncoder(layer, 4, enable_nested_tensor=False).eval()
cama = m.CamaTransformer(enc, groups=10, group_tokens=25, max_batch=16, device=0)
x = torch.randn(b, 250, 1024, device="cuda").bfloat16()
cama.input_view(b).copy_(x)
for _ in range(3):
    cama.forward(b=b, use_graph=False)
torch.cuda.synchronize()
