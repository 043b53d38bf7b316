"""Host model of the 16-bit quotient-tag visited table prepared for the next round
(search_common.cuh Tag16Map / tag16_of, built only with -DDAB_V2_TAG16_BUILD=1): the id ->
(bucket, tag) map must be a bijection on [0, 2^K) with 14-bit tags, and the multiply-shift
division the device uses must be exact."""
import numpy as np
import pytest


def tag_map(K, nbk):
    s = 0
    while (1 << s) < nbk:
        s += 1
    assert K + s <= 32
    magic = ((1 << (K + s)) + nbk - 1) // nbk
    assert magic < (1 << 32)
    return (1 << K) - 1, magic, K + s


@pytest.mark.parametrize("K,nbk", [(20, 283), (20, 64), (20, 577), (17, 16), (22, 300), (22, 1024), (12, 16)])
def test_bucket_tag_is_a_bijection_with_14_bit_tags(K, nbk):
    assert (1 << K) <= nbk << 14, "host picks n_buckets >= 2^K / 2^14"
    kmask, magic, shift = tag_map(K, nbk)
    ids = np.arange(1 << K, dtype=np.uint64)
    h = (ids * np.uint64(0x9E3779B1)) & np.uint64(kmask)          # odd multiplier mod 2^K: a bijection
    assert len(np.unique(h)) == 1 << K
    tag = (h * np.uint64(magic)) >> np.uint64(shift)              # device: 64-bit multiply-shift
    assert np.array_equal(tag, h // np.uint64(nbk)), "multiply-shift division must be exact for h < 2^K"
    bucket = h - tag * np.uint64(nbk)
    assert bucket.max() < nbk and tag.max() < (1 << 14)
    key = bucket * np.uint64(1 << 14) + tag
    assert len(np.unique(key)) == 1 << K, "(bucket, tag) must identify the id"
    # displaced copies carry d = 1, 2 in the top two bits and never look empty
    assert ((2 << 14) | int(tag.max())) < 0xFFFF
