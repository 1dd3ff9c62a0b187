import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    from mathmap_b200 import build
    build.build()


@pytest.fixture(scope="session", autouse=True)
def built_library():
    _ensure_built()


def load_manifest():
    return json.load(open(os.path.join(GOLDEN, "manifest.json")))


def load_png_rgb(name):
    from PIL import Image
    return np.array(Image.open(os.path.join(GOLDEN, "png", name)).convert("RGB"))


def marlene_rgba():
    """tests/marlene.png as the CLI sees it: RGB with alpha forced to 255 (mathmap_cmdline.c:181-183)."""
    rgb = load_png_rgb("marlene.png")
    return np.ascontiguousarray(np.dstack([rgb, np.full(rgb.shape[:2], 255, np.uint8)]))


def filter_source(rel):
    return open(os.path.join(GOLDEN, "filters", rel)).read()


def synthetic_rgba(width, height, seed=1234, alpha=True):
    """Seeded noise blended 50/50 with a smooth gradient."""
    rng = np.random.default_rng(seed)
    noise = rng.integers(0, 256, (height, width, 4), dtype=np.uint8).astype(np.float32)
    yy, xx = np.mgrid[0:height, 0:width].astype(np.float32)
    grad = np.stack([xx / max(1, width - 1) * 255, yy / max(1, height - 1) * 255, (xx + yy) / max(1, width + height - 2) * 255,
                     255 - xx / max(1, width - 1) * 128], axis=2)
    img = (0.5 * noise + 0.5 * grad).astype(np.uint8)
    if not alpha:
        img[:, :, 3] = 255
    return np.ascontiguousarray(img)


def compare_u8(a, b):
    """Returns (percent exact pixels, percent pixels within 1 LSB, max abs diff)."""
    d = np.abs(a.astype(np.int32) - b.astype(np.int32)).max(axis=2)
    return float((d == 0).mean() * 100), float((d <= 1).mean() * 100), int(d.max())
