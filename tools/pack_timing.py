import sys, time
sys.path[:0] = ["napkon-string-matching_b200", "."]
import numpy as np, torch
from napkon_string_matching import synthetic as syn
from napkon_string_matching.gpu import pack, device_pack as dp
from napkon_string_matching.gpu.engine import Engine
eng = Engine(0)
for n in (50_000, 1_000_000):
    pl, f = syn.term_level_sets(n, 9)
    t0 = time.time(); rank = pack.frequency_rank([f], 20000); hp = pack.pack_part_id_sets(pl, f, 20000, rank); t_host = time.time() - t0
    raw = dp.raw_from_parts(pl, f)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        (c,) = eng.device_packer.pack([raw], 20000)
        torch.cuda.synchronize(); t_dev = time.time() - t0
    t0 = time.time(); raw = dp.raw_from_parts(pl, f); t_raw = time.time() - t0
    print(f"term n={n}: host pack {t_host*1e3:.0f} ms, device pack {t_dev*1e3:.1f} ms (+ raw CSR on host {t_raw*1e3:.1f} ms), packed bytes {c.h2d_bytes}")
    lens, flat = syn.token_id_level_sets(n, 5)
    t0 = time.time(); hp = pack.pack_suffix_id_sets(lens, flat, 30000); t_host = time.time() - t0
    raw = dp.raw_from_id_lists(lens, flat)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        (c,) = eng.device_packer.pack([raw], 30000, rank=None)
        torch.cuda.synchronize(); t_dev = time.time() - t0
    print(f"tokenids n={n}: host pack {t_host*1e3:.0f} ms, device pack {t_dev*1e3:.1f} ms")
# strings: default_process + pack_strings on the host against DeviceStringPacker (host join / encode / numpy
# per level included in the device figure; the kernels are two of its stages)
from napkon_string_matching.text.process import default_process
vocab = syn.vocabulary()
for n in (20_000, 200_000):
    left = [[s] for s in syn.question_strings(n, syn.SEED_LEFT, vocab)]
    right = [[s] for s in syn.question_strings(n, syn.SEED_RIGHT, vocab)]
    t0 = time.time()
    hp = pack.pack_strings([[default_process(x) for x in lv] for lv in left], [[default_process(x) for x in lv] for lv in right])
    t_host = time.time() - t0
    sp = dp.DeviceStringPacker(eng)
    for rep in range(3):
        torch.cuda.synchronize(); t0 = time.time()
        dl, dr = sp.pack([left, right])
        torch.cuda.synchronize(); t_dev = time.time() - t0
    print(f"strings 2 x n={n}: host default_process + pack_strings {t_host*1e3:.0f} ms, DeviceStringPacker.pack {t_dev*1e3:.0f} ms")
