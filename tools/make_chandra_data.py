"""Convert the public Chandra HETG facet table and ACIS corner table that the
reference ships as data (NOT code) into the compact numeric files under
``marxs_b200/missions/chandra/data``.

Sources (read from the reference mount, build container only):
  * marxs/missions/chandra/HESSdesign.rdb — "Facet location data for the 336
    HETG facets", MIT HETG group, http://space.mit.edu/HETG/hess/basic.html
  * marxs/missions/chandra/data.py PIX_CORNER_LSI_PAR — Tables 14-17 of
    "ASC Coordinates" (Chandra coordinate memo), LSI chip-corner coordinates.

Run:  python tools/make_chandra_data.py
"""
import os
import re
import sys

import numpy as np

REF = os.environ.get('MARXS_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'marxs_b200',
                   'missions', 'chandra', 'data')

HESS_COLS = ['xc', 'yc', 'zc', 'xu', 'yu', 'zu', 'xuyf', 'yuyf', 'zuyf',
             'xuxf', 'yuxf', 'zuxf', 'xul', 'yul', 'zul', 'xud', 'yud', 'zud']


def main():
    lines = [l.rstrip('\n') for l in open(os.path.join(REF, 'marxs/missions/chandra/HESSdesign.rdb'))
             if l.strip() and not l.startswith('#')]
    names = lines[0].split('\t')
    rows = [l.split('\t') for l in lines[2:]]
    idx = [names.index(c) for c in HESS_COLS]
    with open(os.path.join(OUT, 'hess_facets.csv'), 'w') as f:
        f.write('# Chandra HETG HESS facet design table (336 facets): MIT HETG group,\n')
        f.write('# http://space.mit.edu/HETG/hess/basic.html ; columns as in HESSdesign.rdb\n')
        f.write('hessloc,' + ','.join(HESS_COLS) + '\n')
        for r in rows:
            f.write(r[0] + ',' + ','.join(r[i] for i in idx) + '\n')
    src = open(os.path.join(REF, 'marxs/missions/chandra/data.py')).read()
    with open(os.path.join(OUT, 'acis_corners_lsi.csv'), 'w') as f:
        f.write('# ACIS chip corner LSI coordinates [mm], Tables 14-17 "ASC Coordinates" (1996)\n')
        f.write('chip,corner,x,y,z\n')
        for m in re.finditer(r'ACIS-([IS]\d)-(LL|LR|UR|UL),s,h,"\(([^)]*)\)"', src):
            x, y, z = m.group(3).split()
            f.write('{0},{1},{2},{3},{4}\n'.format(m.group(1), m.group(2), x, y, z))
    print('wrote', OUT)


if __name__ == '__main__':
    sys.exit(main())
