"""Debug helper (GPU box): the CLI's StackRegistrations against the reference's on the case of
tests/test_gpu_cli.py::test_cli_setup_and_stack_registration_match_the_reference_irtk."""
import sys, os, numpy as np, tempfile, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'tests')); sys.path.insert(0, ROOT)
from oracle import ref_irtk as ri
from fetalreconstruction_b200.geometry import rigid_matrix
from fetalreconstruction_b200.phantom import make_dataset, small_config
from fetalreconstruction_b200 import rreg
from fetalreconstruction_b200.reconstruction import Reconstruction
from test_host_cli import write_nifti
tmp = tempfile.mkdtemp()
cfg = small_config(seed=11, vol=56, n_stacks=3, slices=22, size=52, inplane=1.1, spacing=2.0)
cfg.motion_mm = cfg.motion_deg = 0.0; cfg.noise = 3.0; cfg.corrupt_fraction = 0.0; cfg.mask_semi_axis = 0.34
ds = make_dataset(cfg); n = cfg.slices_per_stack
offsets = [np.zeros(6), np.array([2.0, -1.5, 1.0, 2.0, -1.0, 3.0]), np.array([-1.0, 2.5, -2.0, -3.0, 2.0, 1.0])]
names = []
for s, attr in enumerate(ds.stack_attrs):
    vol = np.where(ds.slices[s*n:(s+1)*n] < 0, 0.0, ds.slices[s*n:(s+1)*n]).astype(np.float32)
    p = os.path.join(tmp, f"stack_{s}.nii"); write_nifti(p, vol, rigid_matrix(*offsets[s]) @ attr.image_to_world(), (attr.dx, attr.dy, attr.dz)); names.append(p)
mp = os.path.join(tmp, "mask.nii"); write_nifti(mp, ds.mask.astype(np.float32), ds.vol_attr.image_to_world(), (cfg.vol_voxel,)*3)
dump = os.path.join(tmp, 'dump'); os.mkdir(dump)
r = subprocess.run([os.path.join(ROOT, 'host/SVRreconstructionGPU'), '-o', 'recon.nii.gz', '-i'] + names + ['-m', mp, '--resolution', '1.0', '--smooth_mask', '0', '--dump_setup', dump, '--no_log', '1'], capture_output=True, text=True, cwd=tmp)
print('cli rc', r.returncode)
print('\n'.join(l for l in r.stdout.splitlines() if 'registered' in l or 'StackRegistrations' in l))
rr = ri.Reconstruction()
for p in names: rr.add_stack(ri.Image.read(p), np.zeros(6), 2*cfg.spacing)
mask = ri.Image.read(mp)
rr.call("crop_stack_to_mask", 0, mask.h); rr.create_template(0, 1.0); rr.set_mask(mask, 0.0)
tmpl = rr.stack(0); m = rr.mask()
ta = tmpl.attrs.copy(); td = np.trunc(tmpl.data)
i2w, _ = tmpl.matrices(); _, mw2i = m.matrices(); md = m.data
zz, yy, xx = np.meshgrid(np.arange(td.shape[0]), np.arange(td.shape[1]), np.arange(td.shape[2]), indexing='ij')
P = np.stack([xx, yy, zz, np.ones_like(xx)], -1).reshape(-1, 4).astype(float)
M = (P @ i2w.T) @ mw2i.T
rI = np.where(M[:, :3] > 0, np.floor(M[:, :3] + 0.5), np.ceil(M[:, :3] - 0.5)).astype(int)
inb = (rI[:, 0] >= 0) & (rI[:, 0] < md.shape[2]) & (rI[:, 1] >= 0) & (rI[:, 1] < md.shape[1]) & (rI[:, 2] >= 0) & (rI[:, 2] < md.shape[0])
keep = np.zeros(len(rI), bool); keep[inb] = md[rI[inb, 2], rI[inb, 1], rI[inb, 0]] != 0
tdm = np.where(keep.reshape(td.shape), td, 0.0)
mo = np.eye(4); mo[:3, 3] = ta[6:9]; ta2 = ta.copy(); ta2[6:9] = 0
target = ri.Image.new(ta2, tdm)
b = Reconstruction(0)
for i in (1, 2):
    start = ri.rigid_from_matrix(mo)
    src = rr.stack(i)
    want = ri.rigid_register(target, src, 0, start)
    got, _, ev = rreg.register(b, [rreg.to_grey(tdm), rreg.to_grey(src.data)], [ta2, src.attrs], [0], [1], 0, [start])
    print('stack', i, 'start', np.round(start, 4), '\n   reference engine', want, '\n   device engine   ', got[0], 'evals', ev)
src = rr.stack(1)
start = ri.rigid_from_matrix(mo)
for level in (2, 1, 0):
    ref_sim, rt, rs = ri.reg_probe(target, src, 0, level, start)
    sim, pt, pta, ps, psa = rreg.register(b, [rreg.to_grey(tdm), rreg.to_grey(src.data)], [ta2, src.attrs], [0], [1], 0, [start], level_only=level, want_prepared=True)
    print('level', level, 'sim ref/ours', ref_sim, sim[0], 'attrs equal', np.array_equal(pta, rt.attrs), np.array_equal(psa, rs.attrs),
          'target voxels differing', int((pt != rt.data.astype(np.int16)).sum()) if pt.shape == rt.data.shape else 'shape', 'source voxels differing',
          int((ps != rs.data.astype(np.int16)).sum()) if ps.shape == rs.data.shape else ('shape', ps.shape, rs.data.shape))
    if ps.shape == rs.data.shape and (ps != rs.data.astype(np.int16)).any():
        w = np.argwhere(ps != rs.data.astype(np.int16))[:5]
        print('   first source diffs (z,y,x):', w.tolist(), [int(ps[tuple(i)]) for i in w], [int(rs.data[tuple(i)]) for i in w])
both, _, _ = rreg.register(b, [rreg.to_grey(tdm), rreg.to_grey(rr.stack(1).data), rreg.to_grey(rr.stack(2).data)], [ta2, rr.stack(1).attrs, rr.stack(2).attrs], [0, 0], [1, 2], 0,
                           [ri.rigid_from_matrix(mo)] * 2)
print('device engine, both items in one call', both)
rr.call("stack_registrations", 0)
print('reference StackRegistrations:', [rr.stack_dof(i) for i in (1, 2)])
