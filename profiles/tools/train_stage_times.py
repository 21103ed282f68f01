import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "pytorch-tacotron2-implementation_b200")): sys.path.insert(0, p)
import torch
from b200taco import synth, train as T
from b200taco.model import Tacotron2
B, S, Tn = 64, 800, 150
dev = torch.device("cuda:0")
torch.manual_seed(1234)
m = Tacotron2(); m._projection_bias_initialized = True; m = m.to(dev).train()
eng = m._get_engine()
named = dict(m.named_parameters())
ids = synth.phoneme_ids(B, Tn, seed=0).to(dev)
mel = (torch.randn(B, 80, S) * 2 - 5).to(dev)
lens = torch.full((B,), Tn, dtype=torch.long, device=dev)
def tm(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); out = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, out
for rep in range(2):
    t_enc, memory = tm(lambda: T.EncoderTrainFn.apply(eng, ids, *[named[k] for k in T.encoder_param_keys()]))
    t_dec, (coarse, gate, align) = tm(lambda: T.DecoderTrainFn.apply(eng, memory, mel, lens, T.draw_decoder_masks(B, S, dev), *[named[k] for k in T.decoder_param_keys()]))
    t_post, post = tm(lambda: T.PostnetTrainFn.apply(eng, coarse, T.draw_postnet_masks(B, S, dev), *[named[k] for k in T.postnet_param_keys()]))
    gp = torch.randn_like(post)
    t_bpost, gc = tm(lambda: torch.autograd.grad(post, [coarse] + [named[k] for k in T.postnet_param_keys()], gp, retain_graph=True)[0])
    t_bdec, gm = tm(lambda: torch.autograd.grad([coarse, gate], [memory] + [named[k] for k in T.decoder_param_keys()], [gc, torch.randn_like(gate)], retain_graph=True)[0])
    t_benc, _ = tm(lambda: torch.autograd.grad(memory, [named[k] for k in T.encoder_param_keys()], gm))
    print(dict(enc_fwd=round(t_enc,1), dec_fwd=round(t_dec,1), post_fwd=round(t_post,1), post_bwd=round(t_bpost,1), dec_bwd=round(t_bdec,1), enc_bwd=round(t_benc,1)))
