"""Convert the Si transmission table the reference ships as DATA (Henke optical constants,
http://henke.lbl.gov/optical_constants/filter2.html : Si, density 2.33, thickness 1 micron) into
``marxs_b200/missions/mitsnl/data/si_transmission_1um.csv`` (energy in keV, transmission).

Run in the build container:  python tools/make_mitsnl_data.py"""
import os

REF = os.environ.get('MARXS_REFERENCE_ROOT', '/root/reference')
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'marxs_b200', 'missions', 'mitsnl', 'data')


def main():
    rows = []
    for line in open(os.path.join(REF, 'marxs/missions/mitsnl/data/SiTransmission.csv')):
        line = line.strip()
        if not line or line.startswith('#') or line.startswith('energy'):
            continue
        e, t = line.split()
        rows.append((float(e), float(t)))
    with open(os.path.join(OUT, 'si_transmission_1um.csv'), 'w') as f:
        f.write('# Transmission of 1 micron of Si (density 2.33): Henke et al., henke.lbl.gov/optical_constants/filter2.html\n')
        f.write('energy_eV,transmission\n')
        for e, t in rows:
            f.write('{0!r},{1!r}\n'.format(e, t))
    print('wrote', len(rows), 'rows')


if __name__ == '__main__':
    main()
