"""One-off parity campaign (not collected by pytest): the randomised tests of test_gpu_parity.py /
test_gpu_network.py over many more seeds than the suite runs.

    python tests/fuzz_campaign.py [first_grid_seed n_grid first_net_seed n_net]

Round 1 (B200): 12 000 random single-radar geometries x 3 blind modes + CR and 3 000 random networks —
zero index / mask / value mismatches against the oracle (4.4 minutes of box time).  Round 2, on the r3 / tile
kernels with fresh seeds (`20424 12000 2048 3000`): again zero failures (`profiles/r02_fuzz_campaign.txt`).
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
import test_gpu_parity as T  # noqa: E402
import test_gpu_network as N  # noqa: E402


def campaign(name, fn, first, count):
    bad, t0 = 0, time.time()
    for seed in range(first, first + count):
        try:
            fn(seed)
        except AssertionError as e:
            bad += 1
            print("%s FAIL seed %d: %s" % (name, seed, str(e)[:400].replace("\n", " | ")), flush=True)
        except Exception as e:  # noqa: BLE001
            bad += 1
            print("%s ERROR seed %d: %r" % (name, seed, e), flush=True)
    print("%s: %d seeds, %d failures, %.0f s" % (name, count, bad, time.time() - t0), flush=True)
    return bad


if __name__ == "__main__":
    a = [int(v) for v in sys.argv[1:5]] + [424, 12000, 48, 3000][len(sys.argv[1:5]):]
    bad = campaign("grid fuzz", T.test_randomised_geometry_parity, a[0], a[1])
    bad += campaign("network fuzz", N.test_randomised_network_parity, a[2], a[3])
    sys.exit(1 if bad else 0)
