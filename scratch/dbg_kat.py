import sys; sys.path[:0]=['.','tests']
import numpy as np, kat, orc
from pupiloptixlab_b200 import pb2
import test_gpu_kat as T
F=np.float32
pb2.init(0)
ref=kat.run(orc.port())
WO, WI, WH, xi = ref["ggx_wo"], ref["ggx_wi"], ref["ggx_wh"], ref["ggx_xi"]
for a, al in enumerate(ref["ggx_alpha"]):
    inp = np.concatenate([WI, WO, WH, np.full((len(WO), 1), al, F), xi], 1).astype(F)
    out = np.zeros((len(WO), 8), F)
    pb2.kat("ggx", inp, None, None, len(WO), out)
    r=ref["ggx_sample"][a]
    bad=np.flatnonzero(~np.isclose(out[:,4:7],r,rtol=2e-5,atol=5e-6,equal_nan=True).all(1))
    for i in bad: print("ggx alpha",al,"wo",WO[i],"xi",xi[i],"gpu",out[i,4:7],"ref",r[i])
mats = kat.local_bsdfs()
WO, WI, rng_in = ref["bsdf_wo"], ref["bsdf_wi"], ref["bsdf_rng_in"]
n=len(WO)
for m,b in enumerate(mats):
    arr = (pb2.KatBsdf * n)(*[T._kat_bsdf(b)] * n)
    inp = np.zeros((n, 8), F); inp[:, 0:3], inp[:, 3:6] = WO, WI; inp[:, 6] = rng_in.view(F)
    out = np.zeros((n, 16), F)
    pb2.kat("bsdf", arr, inp, None, n, out)
    s_ref, e_ref = ref["bsdf_sample"][m], ref["bsdf_eval"][m]
    def rep(name,a,r,rtol,atol):
        a64,r64=a.astype(np.float64),r.astype(np.float64)
        ok=np.isclose(a64,r64,rtol=rtol,atol=atol,equal_nan=True)|(np.isinf(a64)&np.isinf(r64))
        bad=np.flatnonzero(~ok.all(1))
        for i in bad[:6]: print("mat",m,"type",b.type,"alpha",b.alpha,name,"row",i,"wo",WO[i],"wi",WI[i],"gpu",a[i],"ref",r[i])
        if len(bad): print("   total bad",len(bad))
    rep("sample_wi",out[:,0:3],s_ref[:,0:3],2e-5,5e-6)
    rep("sample_fpdf",out[:,3:7],s_ref[:,3:7],2e-4,1e-6)
    rep("eval",out[:,9:13],e_ref,2e-4,1e-6)
    if not np.array_equal(out[:, 8].view(np.uint32), ref["bsdf_sample_rng"][m]): print("mat",m,"rng differs")
    if not np.array_equal(out[:, 7].view(np.uint32), ref["bsdf_sample_type"][m]): print("mat",m,"lobe differs")
