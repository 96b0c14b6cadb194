"""Sanity script (GPU box): 400 fused train steps on one fixed synthetic batch (128x256, batch 4, 9 blocks).
The generator loss must fall (the L1 term dominates: 54 -> ~10 when over-fitting one batch) and every weight
must stay finite -- catches races between the main and the side stream that a 3-step parity test could miss."""
import importlib, sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import sggan_oracle as O
L = importlib.import_module("sg-gan-tf2_b200._lib")
B,H,W,C,nb = 4,128,256,34,9
cfg = L.default_config(B,H,W,n_blocks=nb,segment_class=C)
eng = L.Engine(cfg,"cuda:0")
gw = O.init_weights(O.generator_spec(n_blocks=nb),1); dw = O.init_weights(O.discriminator_spec(segment_class=C),2)
eng.set_weights(L.NET_G,gw); eng.set_weights(L.NET_D,dw); eng.weights_changed()
a,s,m,_ = O.synthetic_batch(B,H,W,C,seed=3)
a,s,m = a.cuda(), s.cuda(), m.cuda()
hist=[]
for it in range(400):
    eng.train_step(a,s,m)
    if it % 40 == 0 or it == 399:
        l = eng.losses.cpu().tolist(); hist.append(l); print(it, l, flush=True)
ok = all(np.isfinite(x).all() for x in hist) and hist[-1][0] < hist[0][0]
w = eng.tensors(L.NET_G,0); print("weights finite:", all(torch.isfinite(t).all().item() for t in w), "gen loss decreased:", hist[-1][0] < hist[0][0])
sys.exit(0 if ok else 1)
