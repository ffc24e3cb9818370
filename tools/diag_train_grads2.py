import os, sys, copy, warnings
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
warnings.filterwarnings("ignore")
import cova_b200.synth as synth
from cova_b200.models import CoVA
from conftest import load_golden
DEV = "cuda:0"
g = load_golden("g_train_r18_img128")
for path in ("native", "torch"):
    os.environ["COVA_B200_TRAIN_BACKBONE"] = path
    m = CoVA((3, 3), 128, 4, True, 384, 32, 0, 0.0, None, pretrained=False)
    m.load_state_dict(synth.make_state_dict(123), strict=True)
    m = m.to(DEV).train()
    inp = [t.to(DEV) for t in synth.gen(2, 12, 8, seed=8, img=128, with_labels=True)]
    with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
        out = m(*inp[:4]); loss = torch.nn.CrossEntropyLoss(reduction="sum")(out, inp[4]); loss.backward()
    print(path, "logits err", float(np.abs(out.detach().cpu().numpy() - g["logits"]).max() / np.abs(g["logits"]).max()), "loss", float(loss), float(g["loss"]))
    grads = dict(m.named_parameters())
    for k in [k for k in g if k.startswith("grad:")]:
        got, want = grads[k[5:]].grad.cpu().numpy(), g[k]
        print(f"   {k:40s} err {np.abs(got - want).max() / np.abs(want).max():9.2e}")
    sd = m.state_dict()
    for k in [k for k in g if k.startswith("buf:")][:4]:
        print(f"   {k:40s} err {np.abs(sd[k[4:]].cpu().numpy() - g[k]).max() / np.abs(g[k]).max():9.2e}")
