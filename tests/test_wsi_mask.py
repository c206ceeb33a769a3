"""Row f4: known-answer tests for the tissue mask (`simple_get_mask`, reference infer/wsi.py:489-499).
scikit-image is not installed here, so the reference's three `skimage.morphology` calls are restated from their
documented semantics by brute force inside this test (pure-Python flood fill, literal disk), independently of the
scipy / cv2 implementation under test:
  remove_small_objects(mask, min_size=256, connectivity=2): drop 8-connected components with FEWER than 256 px;
  remove_small_holes(mask, area_threshold=16384)          : fill 4-connected background components with FEWER than
                                                             16384 px (touching the border or not);
  binary_dilation(mask, disk(16))                          : OR over offsets with dy^2 + dx^2 <= 16^2, outside = 0."""
import numpy as np

from hover_net_b200.infer import wsi


def _components(mask, eight):
    lab = np.zeros(mask.shape, np.int32)
    H, W = mask.shape
    nb = [(-1, 0), (1, 0), (0, -1), (0, 1)] + ([(-1, -1), (-1, 1), (1, -1), (1, 1)] if eight else [])
    n = 0
    for y0, x0 in zip(*np.nonzero(mask)):
        if lab[y0, x0]:
            continue
        n += 1
        lab[y0, x0] = n
        stack = [(y0, x0)]
        while stack:
            y, x = stack.pop()
            for dy, dx in nb:
                yy, xx = y + dy, x + dx
                if 0 <= yy < H and 0 <= xx < W and mask[yy, xx] and not lab[yy, xx]:
                    lab[yy, xx] = n
                    stack.append((yy, xx))
    return lab, n


def _expected(dark):
    m = dark.copy()
    lab, n = _components(m, True)
    for k in range(1, n + 1):
        if (lab == k).sum() < 16 * 16:
            m[lab == k] = False
    lab, n = _components(~m, False)
    for k in range(1, n + 1):
        if (lab == k).sum() < 128 * 128:
            m[lab == k] = True
    out = np.zeros_like(m)
    H, W = m.shape
    ys, xs = np.nonzero(m)
    for dy in range(-16, 17):
        for dx in range(-16, 17):
            if dy * dy + dx * dx <= 256:
                yy, xx = ys + dy, xs + dx
                ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
                out[yy[ok], xx[ok]] = True
    return out


def _thumb(dark):
    g = np.where(dark, 60, 245).astype(np.uint8)
    return np.stack([g, g, g], -1)


def test_mask_known_answers():
    H, W = 420, 520
    dark = np.zeros((H, W), bool)
    dark[30:46, 30:46] = True          # 16x16 = 256 px: kept (not FEWER than 256)
    dark[30:45, 80:97] = True          # 15x17 = 255 px: removed
    dark[100:108, 30:46] = True        # 128 px ...
    dark[108:116, 46:62] = True        # ... + 128 px touching only diagonally: one 8-connected object of 256 -> kept
    dark[150:400, 150:450] = True      # tissue block with holes
    dark[160:287, 160:289] = False     # hole 127 x 129 = 16383 px: filled
    dark[160:288, 300:428] = False     # hole 128 x 128 = 16384 px: stays
    dark[300:310, 160:170] = False     # two small holes touching only diagonally: 4-connectivity keeps them separate,
    dark[310:320, 170:180] = False     # both filled
    got = wsi.simple_get_mask(_thumb(dark))
    exp = _expected(dark)
    assert got.shape == exp.shape and np.array_equal(got.astype(bool), exp)
    # spot checks of the rules themselves
    assert got[38, 38] and not got[37, 88 + 30]                      # kept speck / removed speck far from anything
    assert not got[37, 88]                                            # the 255 px speck left nothing behind
    assert got[220, 220] and not got[224, 364]                        # small hole filled, threshold-sized hole kept
    assert got[30 - 16, 38] and not got[30 - 17, 38]                  # dilation reaches exactly 16 px straight up
    assert got[30 - 10, 30 - 12] and not got[30 - 11, 30 - 12]        # 10^2 + 12^2 = 244 <= 256 < 11^2 + 12^2


def test_mask_border_background_counts_as_hole():
    """`remove_small_holes` knows nothing about borders: a background pocket FEWER than 16384 px is filled even when it
    touches the image edge."""
    dark = np.ones((200, 300), bool)
    dark[:100, :100] = False           # 10000 px pocket in the corner
    dark[120:, 150:] = False           # 80 x 150 = 12000 px pocket on two edges
    got = wsi.simple_get_mask(_thumb(dark))
    assert got.all() and np.array_equal(got.astype(bool), _expected(dark))


def test_tiny_thumbnail_quirk():
    """A thumbnail smaller than 128 x 128 px: once the speck is gone the whole image is one background component of
    FEWER than 16384 px, i.e. a "small hole" -- the reference's pipeline then marks everything as tissue."""
    dark = np.zeros((64, 64), bool)
    dark[10:20, 10:20] = True
    got = wsi.simple_get_mask(_thumb(dark))
    assert got.all() and np.array_equal(got.astype(bool), _expected(dark))
