"""Development aid (CPU): how loose is the rigorous a-posteriori bound  ||dC||_inf <= ||A^-1||_inf * ||r||_inf  for the
confidence-map systems?  A is an M-matrix (A^-1 >= 0 entrywise), so ||A^-1||_inf = max(A^-1 1): one more direct solve.
Output kept in profiles/r2_rigorous_bound_study.log; conclusion in DESIGN.md section 4.

    PYTHONPATH=. python tests/manual/bound_study.py
"""
import numpy as np, scipy.sparse.linalg as spla, time
from oracle import confmap as ocm, synth as osynth
for (H,W,kind) in [(96,128,'s'),(256,256,'s'),(256,256,'n'),(512,512,'s')]:
    img,lab = osynth.make_batch(1,H,W)
    im = img[0] if kind=='s' else np.random.default_rng(0).random((H,W)).astype(np.float32)
    planes = ocm.edge_weights(im, 2.0, 90.0, 0.05)
    A,b = ocm.system(*planes)
    t=time.time(); lu = spla.splu(A); x = lu.solve(b); u = lu.solve(np.ones_like(b)); t=time.time()-t
    kinf = u.max()
    for rtol in (1e-6,1e-8):
        out = ocm.pcg_reference(A,b,"line",W,rtol=rtol)
        xk = out[0] if isinstance(out,tuple) else out
        r = b - A@xk
        err = np.abs(xk-x).max()
        print(H,W,kind,rtol,"kinf %.3e |b|2 %.3e err %.3e  bound_inf %.3e (x%.1f)  bound_2 %.3e  relres %.2e K=%.2e Kbound=%.2e"%(kinf,np.linalg.norm(b),err,kinf*np.abs(r).max(),kinf*np.abs(r).max()/err,kinf*np.linalg.norm(r),np.linalg.norm(r)/np.linalg.norm(b),err/(np.linalg.norm(r)/np.linalg.norm(b)), kinf*np.linalg.norm(b)), "t=%.1f"%t)
