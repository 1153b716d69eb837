#!/usr/bin/env python
"""Throughput of the OTHER BASELINE.json configurations (C1, C3, C4) at their full sizes.

The headline bench (bench.py) is C2; this script is the record that the same engine carries the
remaining configurations at size, with the same accounting (240 B/photon/stage +
8 B per diagnostic column) and size-independent sanity checks.  One JSON line per configuration.

    python tools/bench_configs.py [--scale 1.0] [--steps 5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import marxs_b200 as mb  # noqa: E402
from marxs_b200 import optics, simulator, source, _lib  # noqa: E402
from marxs_b200.missions import mitsnl  # noqa: E402
from marxs_b200.program import Lowering  # noqa: E402

PEAK = 6456.2
if os.path.exists(os.path.join(ROOT, 'MEASURED_PEAKS.json')):
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])


def timed(fn, steps):
    fn()
    fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def run_split(name, elements, split, table, n, diag_cols, steps, checks):
    """Experiment: the same trace as TWO launches (elements[:split] out of place, elements[split:] in place on
    the result).  Costs one more pass of the photon record through HBM but halves the code each kernel carries."""
    import ctypes
    base = mb.PhotonBatch(table, device='cuda')
    out = base.copy()
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    status = torch.zeros(_lib.MXB_STATUS_WORDS, dtype=torch.int64, device='cuda')
    stages = []
    for k, els in enumerate((elements[:split], elements[split:])):
        lw = Lowering(out.colnames, meta=out.meta)
        simulator.Sequence(elements=els)._lower(lw)
        prog = lw.finish()
        cols, _ = prog.columns_struct(out)
        stages.append((prog, cols, prog.device_blob(base.device), prog.source_planes(base, n) if k == 0 else None))
    kk = [0]
    infos = []

    def step():
        kk[0] += 1
        for prog, cols, blob, src in stages:
            if src is not None:
                rc = lib.mxb_trace_from(blob.data_ptr(), prog.blob.size, prog.blob.ctypes.data, src, ctypes.byref(cols), n, 0,
                                        99 + kk[0], status.data_ptr(), stream)
            else:
                rc = lib.mxb_trace(blob.data_ptr(), prog.blob.size, prog.blob.ctypes.data, ctypes.byref(cols), n, 0,
                                   99 + kk[0], status.data_ptr(), stream)
            assert rc == 0, lib.mxb_last_error()
            if len(infos) < 2:
                infos.append(lib.mxb_jit_info().decode())
    ms = timed(step, steps)
    nb = 240 + 8 * diag_cols
    print(json.dumps(dict(config=name + ' [split into 2 launches at element %d]' % split, photons=n, kernel_ms=ms,
                          photons_per_s=n / ms * 1e3, algorithmic_bytes_per_photon=nb, achieved_gbs=nb * n / ms / 1e6,
                          frac_of_measured_hbm=nb * n / ms / 1e6 / PEAK, kernel=infos, checks=checks(out))), flush=True)


def run_resident(name, elements, table, n, diag_cols, steps, checks):
    """Out-of-place trace of a resident input table (like bench.py)."""
    inst = simulator.Sequence(elements=elements)
    base = mb.PhotonBatch(table, device='cuda')
    lw = Lowering(base.colnames, meta=base.meta)
    inst._lower(lw)
    prog = lw.finish()
    out = base.copy()
    import ctypes
    cols, _ = prog.columns_struct(out)
    src = prog.source_planes(base, n)
    blob = prog.device_blob(base.device)
    status = torch.zeros(_lib.MXB_STATUS_WORDS, dtype=torch.int64, device='cuda')
    lib = _lib.load()
    stream = torch.cuda.current_stream().cuda_stream
    k = [0]

    def step():
        k[0] += 1
        rc = lib.mxb_trace_from(blob.data_ptr(), prog.blob.size, prog.blob.ctypes.data, src, ctypes.byref(cols), n, 0,
                                99 + k[0], status.data_ptr(), stream)
        assert rc == 0, lib.mxb_last_error()
    ms = timed(step, steps)
    nb = 240 + 8 * diag_cols
    line = dict(config=name, photons=n, kernel_ms=ms, photons_per_s=n / ms * 1e3, algorithmic_bytes_per_photon=nb,
                achieved_gbs=nb * n / ms / 1e6, frac_of_measured_hbm=nb * n / ms / 1e6 / PEAK,
                kernel=lib.mxb_jit_info().decode(), ops=prog.n_ops, diagnostic_columns=len(prog.out_f64) + len(prog.out_i64),
                checks=checks(out))
    print(json.dumps(line), flush=True)


def c1(scale, steps):
    n = int(1e5 * max(scale, 1.))
    rng = np.random.default_rng(1)
    off = np.deg2rad(1. / 60.)
    phi, th = rng.uniform(0, 2 * np.pi, n), rng.uniform(0, off, n)
    d = np.zeros((n, 4))
    d[:, 0], d[:, 1], d[:, 2] = -np.cos(th), np.sin(th) * np.cos(phi), np.sin(th) * np.sin(phi)
    pos = np.zeros((n, 4))
    pos[:, 3] = 1.
    table = dict(pos=pos, dir=d, energy=np.ones(n), polarization=np.tile([0., 1., 0., 0.], (n, 1)), probability=np.ones(n))
    elements = [optics.RectangleAperture(position=[100., 0, 0], zoom=2), optics.PerfectLens(focallength=1000., zoom=400),
                optics.FlatDetector(pixsize=0.01, position=[-1000., 0, 0], zoom=1e5)]
    run_resident('C1 aperture -> lens -> detector', elements, table, n, 8, steps,
                 lambda o: dict(on_detector=float(torch.isfinite(o['det_x']).double().mean())))


def c3_elements(seed=3):
    """CAT-grating spectrograph: lens + scatter -> 561 CATL1L2Stack facets on a Rowland torus -> strip of
    16 CCDs; efficiency table 135 x 25 x 28 (synthetic numbers, real shape).  Host only (no GPU needed)."""
    rng = np.random.default_rng(seed)
    wave = np.linspace(0.5, 7.5, 135)
    theta = np.deg2rad(np.linspace(0.2, 4.0, 25))
    orders = np.arange(-20, 8)[::-1]
    # blazed efficiencies: a Gaussian in order around m0 = -2 sin(theta) d / lambda (d = 200 nm) + some zero order
    m0 = -2. * np.sin(theta)[None, :, None] * 200. / wave[:, None, None]
    # amplitude chosen so that the summed probability stays below 1 in every cell (0.25 x 3.01 x 1.1 + 0.033 = 0.86)
    prob = 0.25 * np.exp(-0.5 * ((orders[None, None, :] - m0) / 1.2) ** 2) + 0.03 * (orders == 0)[None, None, :]
    prob = prob * rng.uniform(0.9, 1.1, prob.shape)
    sel = mitsnl.InterpolateEfficiencyTable(wave, theta, prob, orders)
    # facet placement as SURVEY 8(d): 561 facets on RowlandTorus(6000, 6000), blazed by 1.91 deg; 16 CCDs on the
    # Rowland circle (marxs_b200.design.rowland, pinned to the reference's placement in tests/test_rowland.py)
    from marxs_b200.design import RowlandTorus, GratingArrayStructure, RectangularGrid
    from marxs_b200.affines import axangle2mat
    rt = RowlandTorus(6000., 6000.)
    gas = GratingArrayStructure(rowland=rt, d_element=[30., 30.], radius=[300., 500.], elem_class=mitsnl.CATL1L2Stack,
                                elem_args={'zoom': [1, 13.5, 13.5], 'order_selector': sel,
                                           'orientation': axangle2mat([0, 0, 1], np.deg2rad(1.91))})
    det = RectangularGrid(rowland=rt, d_element=[49.652, 49.652], y_range=[-50, 700], elem_class=optics.FlatDetector,
                          elem_args={'zoom': [1, 24.576, 12.288], 'pixsize': 0.024}, id_col='CCD_ID', guess_distance=25.)
    n_facets = len(gas.elements)
    elements = [
        optics.FlatStack(position=[12000., 0, 0], zoom=[1, 2000, 2000], elements=[optics.PerfectLens, optics.RadialMirrorScatter],
                         keywords=[{'focallength': 12000.}, {'inplanescatter': 1e-5, 'perpplanescatter': 1e-6}]),
        gas, det]
    return elements, n_facets


def c3_setup(n, seed=3):
    """The C3 instrument and n photons on the lens, resident on the GPU."""
    elements, n_facets = c3_elements(seed)
    g = torch.Generator(device='cuda').manual_seed(5)
    b = mb.PhotonBatch(device='cuda')
    pos = b.new_column('pos', torch.float64, vector=True, n=n)
    pos[0] = 12100.
    # the facets sit 4-20 mm behind the lens (RowlandTorus(6000, 6000) reaches x = 12000): an annulus [285, 515] mm
    rad = torch.sqrt(285. ** 2 + torch.rand(n, device='cuda', generator=g, dtype=torch.float64) * (515. ** 2 - 285. ** 2))
    phi = torch.rand(n, device='cuda', generator=g, dtype=torch.float64) * (2 * np.pi)
    pos[1] = rad * torch.cos(phi)
    pos[2] = rad * torch.sin(phi)
    pos[3] = 1.
    d = b.new_column('dir', torch.float64, fill=0., vector=True)
    d[0] = -1.
    p = b.new_column('polarization', torch.float64, fill=0., vector=True)
    p[1] = 1.
    b.new_column('energy', torch.float64)[...] = 0.3 + 1.2 * torch.rand(n, device='cuda', generator=g, dtype=torch.float64)
    b.new_column('probability', torch.float64, fill=1.)
    return elements, b, n_facets


def c3(scale, steps, split=0):
    n = int(1e8 * scale)
    elements, b, n_facets = c3_setup(n)
    chk = lambda o: dict(on_facet=float((o['facet'] >= 0).double().mean()),   # noqa: E731
                         on_ccd=float((o['CCD_ID'] >= 0).double().mean()), mean_order=float(torch.nanmean(o['order'])))
    if split:
        run_split('C3 CAT spectrograph', elements, split, b, n, 19, steps, chk)
        return
    run_resident('C3 CAT spectrograph: lens+scatter -> {0} CATL1L2Stack facets -> 16 CCDs'.format(n_facets), elements, b, n,
                 19, steps, lambda o: dict(on_facet=float((o['facet'] >= 0).double().mean()),
                                           on_ccd=float((o['CCD_ID'] >= 0).double().mean()),
                                           mean_order=float(torch.nanmean(o['order']))))


def c4(scale, steps):
    """Multilayer-mirror polarimeter, photons BORN on the device (LabPointSourceCone -> MultiLayerMirror ->
    FlatBrewsterMirror -> FlatDetector): nothing is read, only results are written."""
    n = int(1e8 * scale)
    g = dict(np.load(os.path.join(ROOT, 'tests', 'golden', 'mlmirror.npz')))
    refl = {'X(mm)': g['ml_x_mm'], 'Peak lambda': g['ml_peak_lambda'], 'Peak': g['ml_peak'], 'FWHM(nm)': g['ml_fwhm']}
    polt = {'Photon energy': g['ml_pol_energy_ev'], 'Polarization': g['ml_pol']}
    a = 2 ** -0.5
    rot1 = np.array([[a, 0, -a], [0, 1, 0], [a, 0, a]])
    rot2 = np.array([[0, 0, 1.], [a, a, 0], [-a, a, 0]])
    elements = [optics.MultiLayerMirror(reflFile=refl, testedPolarization=polt, orientation=rot1, zoom=[1, 24.5, 12.]),
                optics.FlatBrewsterMirror(orientation=rot2, position=[0., 0., 30.], zoom=[1, 10., 30.]),
                optics.FlatDetector(pixsize=0.05, position=[0., 50., 30.], orientation=np.array([[0, -1., 0], [1., 0, 0], [0, 0, 1.]]),
                                    zoom=[1, 20., 20.])]
    src = source.LabPointSourceCone(position=[200., 0, 0], direction=[-1., 0, 0], half_opening=0.02, flux=float(n), energy=0.31)
    out = [None]

    def step():
        out[0] = source.observe(src, None, elements, 1., device='cuda', check=False, out=out[0])
    # allocation of the result table is part of observe(); time the launches only through the kernel info
    ms = timed(step, steps)
    o = out[0]
    nb = 120 + 8 * (len(o.colnames) - 5)       # born: the record is only written
    print(json.dumps(dict(config='C4 multilayer polarimeter, born on device (observe)', photons=n, ms_per_observation=ms,
                          photons_per_s=n / ms * 1e3, algorithmic_bytes_per_photon=nb, achieved_gbs=nb * n / ms / 1e6,
                          frac_of_measured_hbm=nb * n / ms / 1e6 / PEAK, kernel=_lib.load().mxb_jit_info().decode(),
                          note='result table reused between observations (observe(..., out=))',
                          checks=dict(on_detector=float(torch.isfinite(o['det_x']).double().mean()),
                                      mean_probability_detected=float(o['probability'][torch.isfinite(o['det_x'])].mean())))),
          flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--scale', type=float, default=1.0, help='fraction of the full photon counts')
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--only', default='')
    ap.add_argument('--split', type=int, default=0, help='experiment (c3): two launches, split before this element')
    a = ap.parse_args()
    for name, fn in (('c1', c1), ('c3', c3), ('c4', c4)):
        if a.only and name not in a.only.split(','):
            continue
        t0 = time.time()
        if name == 'c3' and a.split:
            fn(a.scale, a.steps, a.split)
        else:
            fn(a.scale, a.steps)
        torch.cuda.empty_cache()
        print('# {0} done in {1:.1f} s'.format(name, time.time() - t0), file=sys.stderr)


if __name__ == '__main__':
    main()
