timeout 900 python -m pytest tests/test_loader.py -m gpu -x -q 2>&1 | tail -25
python - <<'PY'
import time, torch
from paintrl_b200 import loader
t = time.time(); p = loader.load_part('tests/data/urdf/painting/bulge.urdf', max_points=5200, part_no=100); print('bulge load (GPU rasteriser) %.2f s' % (time.time() - t), p.meta['loader_stats'])
t = time.time(); p = loader.load_part('tests/data/urdf/painting/bulge.urdf', max_points=5200, part_no=100, texture_size=(2048, 2048)); print('bulge 2048x2048 %.2f s, %d texels' % (time.time() - t, p.n_texels))
PY
