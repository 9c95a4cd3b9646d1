"""Generates the synthetic test part `bulge` (tests/data/urdf/painting/bulge.{urdf,obj,mtl,png}).

The reference's part meshes live only under /root/reference and do not travel to the GPU box; this part is the
repository's own: a bulged panel 0.8 m x 1.0 m with a dent (so both normal-correction passes have work: the dent's
walls are more than 30 degrees off the hull facet above them and more than 10 degrees off their neighbours), a back
face and a rim, front and back UV islands side by side in a 128 x 128 texture whose red channel carries the label
values 0 and 191 here and there (`RGBColorHandler.change_pixel` leaves such texels alone, bullet_paint_wrapper.py:358-365).
With 31 vertex rows under a 100-row silhouette grid most grid rows are sparse (bullet_paint_wrapper.py:934-947).

    python tests/data/make_synthetic_part.py        # rewrites the four files (deterministic)
"""
import math
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, 'urdf', 'painting')
NY, NZ = 24, 30
LY, LZ = 0.8, 1.0
THICK = 0.02


def height(y, z):
    bulge = 0.09 * math.sin(math.pi * y / LY) * math.sin(math.pi * z / LZ)
    dy, dz = (y - 0.5) / 0.07, (z - 0.62) / 0.07
    dent = -0.05 * math.exp(-(dy * dy + dz * dz))
    return bulge + dent


def main():
    os.makedirs(OUT, exist_ok=True)
    v, vt, faces = [], [], []

    def vid(layer, iy, iz):
        return layer * (NY + 1) * (NZ + 1) + iy * (NZ + 1) + iz + 1

    for layer in range(2):
        for iy in range(NY + 1):
            for iz in range(NZ + 1):
                y, z = LY * iy / NY, LZ * iz / NZ
                x = height(y, z) - (THICK if layer else 0.0)
                v.append((x, y, z))
                u0 = 0.03 + 0.45 * iy / NY + (0.5 if layer else 0.0)
                vt.append((u0, 0.04 + 0.92 * iz / NZ))
    for iy in range(NY):
        for iz in range(NZ):
            a, b, c, d = vid(0, iy, iz), vid(0, iy + 1, iz), vid(0, iy + 1, iz + 1), vid(0, iy, iz + 1)
            faces += [(a, b, c), (a, c, d)]                       # counter-clockwise seen from +x
            a, b, c, d = vid(1, iy, iz), vid(1, iy + 1, iz), vid(1, iy + 1, iz + 1), vid(1, iy, iz + 1)
            faces += [(a, c, b), (a, d, c)]                       # back: seen from -x
    rim = [(iy, 0, iy + 1, 0) for iy in range(NY)] + [(NY, iz, NY, iz + 1) for iz in range(NZ)] + \
          [(iy + 1, NZ, iy, NZ) for iy in range(NY)] + [(0, iz + 1, 0, iz) for iz in range(NZ)]
    for y0, z0, y1, z1 in rim:
        a, b, c, d = vid(0, y0, z0), vid(1, y0, z0), vid(1, y1, z1), vid(0, y1, z1)
        faces += [(a, b, c), (a, c, d)]
    with open(os.path.join(OUT, 'bulge.obj'), 'w') as f:
        f.write('# synthetic test part of paintrl_b200 (tests/data/make_synthetic_part.py)\nmtllib bulge.mtl\no bulge\n')
        for p in v:
            f.write('v %.6f %.6f %.6f\n' % p)
        for t in vt:
            f.write('vt %.6f %.6f\n' % t)
        f.write('usemtl panel\ns off\n')
        for tri in faces:
            f.write('f %s\n' % ' '.join('%d/%d' % (k, k) for k in tri))
    with open(os.path.join(OUT, 'bulge.mtl'), 'w') as f:
        f.write('newmtl panel\nKa 1.000000 1.000000 1.000000\nKd 0.640000 0.640000 0.640000\nKs 0.500000 0.500000 0.500000\n'
                'Ns 96.078431\nd 1.000000\nillum 2\nmap_Kd bulge.png\n')
    with open(os.path.join(OUT, 'bulge.urdf'), 'w') as f:
        f.write('''<?xml version="1.0" ?>
<robot name="bulge">
  <link name="base">
    <inertial><origin rpy="0 0 0" xyz="0 0 0"/><mass value="0"/><inertia ixx="0" ixy="0" ixz="0" iyy="0" iyz="0" izz="0"/></inertial>
    <visual><origin rpy="0 0 0" xyz="0 0 0"/><geometry><mesh filename="bulge.obj" scale="1 1 1"/></geometry></visual>
    <collision><origin rpy="0 0 0" xyz="0 0 0"/><geometry><mesh filename="bulge.obj" scale="1 1 1"/></geometry></collision>
  </link>
</robot>
''')
    from PIL import Image
    rng = np.random.default_rng(7)
    img = rng.integers(1, 255, size=(128, 128, 3), dtype=np.uint8)
    img[rng.random((128, 128)) < 0.05, 0] = 0
    img[rng.random((128, 128)) < 0.05, 0] = 191
    img[rng.random((128, 128)) < 0.03, 0] = 255
    Image.fromarray(img, 'RGB').save(os.path.join(OUT, 'bulge.png'), optimize=True)
    print('wrote', OUT, len(v), 'vertices', len(faces), 'faces')


if __name__ == '__main__':
    main()
