import os, sys, torch, numpy as np
sys.path.insert(0, '/root/repo')
os.environ["NZ_RL_MIN_ELTS"]="0"
from tests.test_scan_gpu import _seeded, _oracle
from tests.helpers import rel_err
import nnuzoo_b200.selective_scan_interface as ssi
inp, gout = _seeded(2,128,2,16,4128,False,seed=13+4128)
ref_out, ref_last, ref_g = _oracle(inp, gout)
dev=torch.device('cuda:0')
def run(fwd, nocp=False, items=None):
    os.environ["NZ_RL_FWD"]=fwd
    if nocp: os.environ["NZ_NO_CP"]="1"
    else: os.environ.pop("NZ_NO_CP",None)
    if items: os.environ["NZ_RL_ITEMS"]=items
    else: os.environ.pop("NZ_RL_ITEMS",None)
    l={k:(None if v is None else v.to(dev).requires_grad_(True)) for k,v in inp.items()}
    out,last=ssi.selective_scan_fn(l["u"],l["delta"],l["A"],l["B"],l["C"],l["D"],l["z"],l["delta_bias"],True,True)
    out.backward(gout.to(dev)); torch.cuda.synchronize()
    e={g: rel_err(l[k].grad.float().cpu().numpy(), ref_g[g]) for g,k in (("du","u"),("ddelta","delta"),("dA","A"),("dB","B"),("dC","C"))}
    dc=(l["C"].grad.cpu()-torch.from_numpy(ref_g["dC"]).float()).abs()
    bad=(dc>1e-3*np.abs(ref_g["dC"]).max()).nonzero()
    print(fwd, nocp, items, {k:'%.1e'%v for k,v in e.items()}, 'bad dC t:', torch.unique(bad[:,3]).tolist()[:20], 'n', torch.unique(bad[:,2]).tolist(), 'bg', torch.unique(bad[:,0]*2+bad[:,1]).tolist())
for rep in range(2):
    run("0"); run("1"); run("0", nocp=True); run("0", items="1"); run("0", items="8")
