"""CPU replay of the GPU decoder's entropy stages (same source, compiled by g++) against the oracle.

This is how the kernels' parsers, table builders and bit readers are validated in the GPU-less build container;
the warp-cooperative execute stage and the checksum kernel are covered by the -m gpu tests."""
import random

from tests import helpers


def test_golden_vectors(hostsim, oracle):
    for name, frame, raw in helpers.golden_vectors():
        r, out = hostsim.decompress(frame, len(raw), oracle)
        assert r == len(raw) and out == raw, name


def test_valid_frames_match_oracle(hostsim, oracle):
    for frame, data in helpers.make_frames(201, 200):
        n = len(data)
        for cap in (n, n + 13, max(0, n - 1)):
            ro, oo, _ = oracle.decompress(frame, cap)
            rh, oh = hostsim.decompress(frame, cap, oracle)
            assert ro == rh and oo == oh, (n, cap, hex(ro), hex(rh))


def test_fuzzed_frames_same_verdict(hostsim, oracle):
    rng = random.Random(77)
    frames = helpers.make_frames(202, 120)
    verdict_mismatch = 0
    for frame, data in frames:
        if len(frame) < 12:
            continue
        for _ in range(6):
            b = helpers.mutate(rng, frame)
            cap = len(data) + rng.choice([0, 0, 5])
            ro, oo, over = oracle.decompress(b, cap)
            rh, oh = hostsim.decompress(b, cap, oracle)
            eo, eh = helpers.is_err(ro), helpers.is_err(rh)
            if eo != eh:
                # documented deviation: the reference accepts a final sequence whose bits were read
                # past the stream start; the GPU path reports corruption_detected
                assert over and eh and not eo, (hex(ro), hex(rh))
            elif not eo:
                assert ro == rh and oo == oh
