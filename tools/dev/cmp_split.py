"""Dev check: forward + gradient with the 2-CTA split vs the single-CTA kernel (run each in a fresh process)."""
import os, subprocess, sys, numpy as np
args = sys.argv[1:4]
code = r'''
import sys, numpy as np
sys.path.insert(0, '.')
from hmcmt2d_b200 import api, synthetic
ny, nz, nf = (int(a) for a in sys.argv[1:4])
mesh, data, inv, prior = synthetic.make_problem(ny, nz, nf, nRx=10)
m = synthetic.stress_model(inv)
pl = api.Plan(mesh, data, inv, prior)
pred, phi, g = pl.forward_gradient(m)
np.savez(sys.argv[4], pred=pred, phi=phi, g=g)
print('ok', phi, flush=True)
'''
outs = []
for sp in ('0', '1'):
    out = '/tmp/cmp_split_%s.npz' % sp
    env = dict(os.environ, HMCMT_SPLIT=sp)
    r = subprocess.run(['timeout', '120', sys.executable, '-c', code, *args, out], env=env, capture_output=True, text=True)
    print('split', sp, 'rc', r.returncode, r.stdout.strip()[-200:], r.stderr.strip()[-600:], flush=True)
    outs.append(out if r.returncode == 0 else None)
if all(outs):
    a, b = np.load(outs[0]), np.load(outs[1])
    print('pred rel', float((np.abs(a['pred'] - b['pred']) / np.abs(a['pred'])).max()),
          'phi rel', abs(float(a['phi']) - float(b['phi'])) / abs(float(a['phi'])),
          'g rel', float(np.abs(a['g'] - b['g']).max() / np.abs(a['g']).max()))
