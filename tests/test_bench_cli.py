"""bench.py's contract, checked without a GPU: the reference arm (the CPU restatement of the decoder) prints one JSON
line whose keys and `config` object are the ones the repo arm prints (the driver compares the two arms' configs key by
key), and the corpus builders behind the sub-records produce what their records say."""
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_line(tmp_path):
    env = dict(os.environ, LZMA_B200_CACHE=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--streams", "8",
                        "--distinct", "4", "--size", "65536", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "GB/s" and line["higher_is_better"] is True
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["e2e"]["value"] == line["value"]
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    cfg = line["config"]
    # the repo arm builds its `config` with the same function from the same corpus: same keys, same values
    assert set(cfg) == {"workload", "units_total", "distinct_streams", "compressed_bytes_total", "decompressed_bytes_total", "sharding", "l2"}
    assert cfg["units_total"] == 16 and cfg["decompressed_bytes_total"] == 16 * 65536 and cfg["distinct_streams"] == 4


def test_config5_corpus_is_varied_and_decodable(tmp_path):
    """corpus.varied_block / build_varied_streams: distinct seeds, a spread of compressed sizes for the scheduler to
    balance, and streams the oracle decodes to the recorded CRC."""
    import zlib
    sys.path.insert(0, ROOT)
    from lzma_b200 import corpus as K
    from oracle import oracle as O
    streams, crcs = K.build_varied_streams(24, 1 << 18, seed0=0, workers=2, preset=1)
    assert len({bytes(s) for s in streams}) == 24
    sizes = np.array([len(s) for s in streams])
    assert sizes.max() > 1.3 * sizes.min()
    for s, c in zip(streams[::5], crcs[::5]):
        r = O.lzma_alone(s, 1 << 18)
        assert r.status == O.OK and zlib.crc32(r.data) == c


def test_mixed_batch_has_every_kind():
    sys.path.insert(0, ROOT)
    from lzma_b200 import corpus as K
    items = K.mixed_batch(unit_size=16 << 10, assets_dir=os.path.join(ROOT, "tests", "golden", "ref_assets"))
    kinds = {i["kind"] for i in items}
    names = " ".join(i["name"] for i in items)
    assert kinds == {"alone", "raw", "lzma2"} and len(items) >= 200
    for needle in ("relabel_lc8lp4pb4", "size_no_eos", "noise_", "lzma2_truncated", "corrupt_", "asset_bad_corrupted.lzma", "asset_randomfile.dat.lzma2"):
        assert needle in names, needle
