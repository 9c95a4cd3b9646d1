"""Mint golden texture planes (Part.texels, bullet_paint_wrapper.py:467, 737-738) from the VERBATIM reference.

TEST INFRASTRUCTURE ONLY.  Run in the build container (needs /root/reference):

    python -m oracle.make_texture_golden        # -> tests/golden/t_texture.npz

For the door panel (RGB) and the quadratic sheet (HSI): reset at start point 0, 25 seeded random discrete
steps, then the reference's whole `texels` array next to its front-texel status plane -- what
`PartPack.compose_texture` must rebuild from the status plane alone.
"""
import os
import random

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, 'tests', 'golden', 't_texture.npz')


def run(part_no, color_mode, seed):
    from oracle.ref_env import ReferenceEnv
    random.seed(seed)
    np.random.seed(seed)
    ref = ReferenceEnv({'Part_NO': part_no, 'COLOR_MODE': color_mode, 'OVERLAP_PENALTY': True})
    ref.reset(0)
    rng = np.random.default_rng(seed)
    for _ in range(25):
        _, _, done, _ = ref.step(int(rng.integers(0, 4)))
        if done:
            break
    texels = np.array([int(v) for v in ref.part.texels], dtype=np.int64)
    return ref.front_status().astype(np.int16), texels.astype(np.int16)


def main():
    out = {}
    for name, part_no, mode in (('door_rgb', 0, 'RGB'), ('sheet_hsi', 1, 'HSI')):
        status, texels = run(part_no, mode, 11)
        out[name + '/status'], out[name + '/texels'] = status, texels
        print(name, 'front texels changed:', int((status != status.max()).sum()) if mode == 'HSI' else int((status == 255).sum()),
              'texel range', texels.min(), texels.max())
    np.savez_compressed(OUT, **out)
    print('wrote', OUT)


if __name__ == '__main__':
    main()
