"""Dump GPU-vs-oracle differences for chosen box-room frames (development aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sp_slam_b200 import api, scenes
from oracle import pyoracle as po
from tests.parity import compare_frame, same_f32

frames = [int(a) for a in sys.argv[1:]] or [80]
P = scenes.poses(1000)
seq = scenes.render(scenes.boxroom_rects(), P[frames], scenes.TUM1)
ext = api.PlaneExtractor(debug=True)
np.set_printoptions(precision=9, linewidth=200)
for k, f in enumerate(frames):
    d = seq[k]
    fp = ext.extract(d)
    orc = po.Oracle().run(d)
    mg, mr = ext.models(0), orc.models()
    print('frame', f, 'models', len(mg), len(mr), 'planes', fp.mnRealPlaneNum, fp.mnPlaneNum, orc.n_real, orc.n_planes)
    for i, (a, b) in enumerate(zip(mg, mr)):
        print(' model', i, 'label', a['label'], b['label'], 'nseg', a['n_segment'], b['n_segment'], 'ninl', len(a['inliers']), len(b['inliers']), 'ncont', len(a['contour']), len(b['contour']))
        print('  coef g', a['coef'], '\n  coef r', b['coef'])
        print('  cen  g', a['centroid'], '\n  cen  r', b['centroid'])
        print('  cov  g', a['cov'].ravel(), '\n  cov  r', b['cov'].ravel())
        print('  curv', a['curvature'], b['curvature'])
        if len(a['inliers']) == len(b['inliers']):
            print('  inliers equal', np.array_equal(a['inliers'], b['inliers']), 'contour equal', np.array_equal(a['contour'], b['contour']))
    try:
        print(' compare:', compare_frame(ext, orc, d, fp))
    except AssertionError as e:
        print(' FAIL:', str(e)[:400])
    print(' times', ext.times(), 'launches', ext.launches)
