import sys, os, torch, time
sys.path.insert(0, "/root/repo")
from oracle import fixtures, vocoder as ovoc
from vectorquantizedcpc_b200 import Vocoder
dev = torch.device("cuda:0")
voc = Vocoder(); voc.load_state_dict(ovoc.init_state_dict(seed=13)); voc = voc.to(dev).eval()
for B, Tc in ((1, 500), (64, 500), (130, 100)):
    codes, spk, u = fixtures.vocoder_inputs(B, Tc, seed=0)
    cd, sd, ud = codes.to(dev), spk.to(dev), u.to(dev)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    wav, x = voc.generate(cd, sd, uniforms=ud, return_mulaw=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    short, xs = voc.generate(cd, sd, uniforms=ud[:, :4000], n_steps=4000, return_mulaw=True)
    ok = torch.equal(xs, x[:, :4000]) and bool(torch.isfinite(wav).all())
    hist = torch.bincount(x.flatten(), minlength=256).float()
    print(f"B={B} L={wav.shape[1]}: {dt*1e3:.0f} ms, {B*wav.shape[1]/16000/dt:.1f}x RT, prefix-consistent={ok}, classes used={int((hist>0).sum())}")
