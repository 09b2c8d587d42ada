/* CPU oracle: 3-D connected-component labelling, 6-connectivity.
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py) -- never linked into the product.
 *
 * Restates the behaviour of the third-party call the reference makes at
 * /root/reference/light_unet/models/metrics.py:50,61  (scipy.ndimage.label with
 * the default structuring element = face neighbours only; scipy is an
 * un-vendored dependency, requirements.txt:5 "scipy>=1.10.0", no pin).
 * Published contract restated here: foreground = non-zero; components are
 * numbered 1..n in C-order (raster) of their first voxel.
 * Algorithm: classic two-pass union-find with path halving; the representative
 * of a set is always its smallest linear index, so ranking the roots by index
 * gives the raster numbering.
 *
 * Pinned in tests/test_oracle_golden.py against scipy.ndimage.label itself
 * (present in the image) and against the golden bbox fixtures generated from
 * the reference.
 */
#include <stdint.h>
#include <stdlib.h>

static int64_t find_root(int64_t *parent, int64_t i) {
    while (parent[i] != i) {
        parent[i] = parent[parent[i]];
        i = parent[i];
    }
    return i;
}

static void unite(int64_t *parent, int64_t a, int64_t b) {
    a = find_root(parent, a);
    b = find_root(parent, b);
    if (a == b) return;
    if (a < b) parent[b] = a; else parent[a] = b;
}

/* mask: D*H*W int32 (non-zero = foreground); labels out: D*H*W int32.
 * returns number of components, or -1 on allocation failure. */
int ccl6_label(const int32_t *mask, int32_t *labels, int D, int H, int W) {
    const int64_t n = (int64_t)D * H * W;
    int64_t *parent = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n > 0 ? n : 1));
    if (!parent) return -1;
    const int64_t sH = W, sD = (int64_t)H * W;
    for (int64_t i = 0; i < n; ++i) parent[i] = i;
    for (int z = 0; z < D; ++z)
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                const int64_t i = z * sD + y * sH + x;
                if (!mask[i]) continue;
                if (x > 0 && mask[i - 1]) unite(parent, i, i - 1);
                if (y > 0 && mask[i - sH]) unite(parent, i, i - sH);
                if (z > 0 && mask[i - sD]) unite(parent, i, i - sD);
            }
    int32_t next = 0;
    /* roots have the smallest index of their set, so a raster scan meets every
     * root before any other member: number roots as they appear. */
    for (int64_t i = 0; i < n; ++i) {
        if (!mask[i]) { labels[i] = 0; continue; }
        const int64_t r = find_root(parent, i);
        if (r == i) labels[i] = ++next;
        else labels[i] = labels[r];
    }
    free(parent);
    return next;
}
