"""Differential fuzz of the two scene front-ends (C++ host library vs Python mirror): mutated files must be accepted or rejected
by both, with the same triangle count, and never crash.  usage: tools/fuzz_*_frontends.py [seed] [iterations]"""
import os, random, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
from cutrace_b200 import host
from cutrace_b200.scene import SceneError, load_scene_json
host.load()
src = open(ROOT + "/scenes/solids.json").read()
rng = random.Random(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
d = tempfile.mkdtemp()
tokens = ['{', '}', '[', ']', ',', ':', '"', '1e999', '-', 'null', 'true', '"type"', '"mesh"', '"file"', '1.5', '\\', 'nan', 'NaN', 'Infinity', '[[[[', '"material": 99', '"material": -1', '"material": 1.5', '"width": 0', '"width": 1e12', '"radius": "x"', '0x10', '01', '.5', '5.', '+1', '1e', "'a'", '/*c*/', '\t', 'é', '\\u0041', '"a":1,"a":2']
agree = dis = 0
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 1500):
    s = list(src)
    for _ in range(rng.randint(1, 3)):
        k = rng.randrange(3)
        pos = rng.randrange(len(s))
        if k == 0: del s[pos:pos + rng.randint(1, 12)]
        elif k == 1: s[pos:pos] = list(rng.choice(tokens))
        else: s[pos] = rng.choice('{}[],:"0123456789.-e ')
    p = os.path.join(d, "f.json")
    open(p, "w").write("".join(s))
    a = b = None; ea = eb = ""
    try: a = host.load_scene(p, base_dir=ROOT)
    except SceneError as e: ea = str(e)
    try: b = load_scene_json(p, base_dir=ROOT)
    except SceneError as e: eb = str(e)
    same = (a is None) == (b is None)
    if same and a is not None:
        da, db = a.to_npz_dict(), b.to_npz_dict()
        same = all(np.array_equal(np.asarray(da[k]), np.asarray(db[k]), equal_nan=True) for k in da)
    if same: agree += 1
    else:
        dis += 1
        if dis <= 12:
            print("DISAGREE cpp:", "ok" if a is not None else ea[:90], "| py:", "ok" if b is not None else eb[:90])
            open(f"/tmp/cutrace_jmism{dis}.json", "w").write("".join(s))
print("agree", agree, "disagree", dis)
