"""Diagnostic: which images / calls of the trunk differ between identical invocations."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tools.gpu_probe import setup_net

net = setup_net(tempfile.mkdtemp())
g = torch.Generator(device="cpu").manual_seed(70)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 70
x0 = torch.randn(B, 3, 224, 224, generator=g).cuda()
x1 = torch.randn(B, 3, 224, 224, generator=g).cuda()
def diff(a, b, tag):
    d = (a != b)
    rows = d.any(1).nonzero().flatten().tolist()
    print("%s: %d elements differ, max abs %.3e, images %s" % (tag, int(d.sum()), (a - b).abs().max().item(), rows[:40]), flush=True)
a1 = net.forward_feat_ext_pair(x0, x1).clone()
a2 = net.forward_feat_ext_pair(x0, x1).clone()
diff(a1, a2, "pair #1 vs #2 (back to back)")
b = net.forward_feat_ext(torch.cat([x0, x1])).clone()
a3 = net.forward_feat_ext_pair(x0, x1).clone()
diff(a1, a3, "pair #1 vs #3 (single-tensor call in between)")
diff(a2, a3, "pair #2 vs #3")
b2 = net.forward_feat_ext(torch.cat([x0, x1])).clone()
diff(b, b2, "single #1 vs #2")
for n in (64, 12, 6):
    y = x0[:n].contiguous()
    c1 = net.forward_feat_ext(y).clone(); c2 = net.forward_feat_ext(y).clone()
    diff(c1, c2, "n=%d twice" % n)
    net.forward_feat_ext(x1[:7].contiguous())
    c3 = net.forward_feat_ext(y).clone()
    diff(c1, c3, "n=%d after another call" % n)
