"""Golden fixture for the duration smoothing of inference.py:248-257, from the UNMODIFIED reference statements.
Authoring container only (needs /root/reference):

    python tests/golden/make_golden_smooth.py

inference.py cannot be imported here (librosa, noisereduce, nltk are not installed), and the smoothing is inline in
`StyleTTS2.__inference`, so this script lifts the statements out of the reference FILE with `ast` and executes them as they are:
the `if prev_d_mean != 0: ... else: ...` block through `pred_dur = torch.round(...)` (inference.py:248-257) and the method
`__replace_outliers_zscore` (inference.py:134-148).  Nothing of the reference is copied into the repository.

`normal_` draws from torch's global generator: every case seeds it, runs the reference statements, then re-seeds and draws
`torch.empty(shape).normal_(0, 1)` -- the N(0, 1) tape the same call consumed (normal_(mean, std) is that draw times std plus
mean) -- which is what st2_smooth_durations takes as `noise`.

Fixture smooth_cases.npz: per case k  duration_k [1, L] (input), noise_k [1, L], t_k, speed_k, prev_k, out_k [1, L] (after
inference.py:255), pred_k [L] (inference.py:257), mean_k (inference.py:272).
"""
import ast
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF = "/root/reference/inference.py"


def lift():
    tree = ast.parse(open(REF).read())
    cls = next(n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "StyleTTS2")
    meth = {n.name: n for n in cls.body if isinstance(n, ast.FunctionDef)}
    outl = meth["__replace_outliers_zscore"]
    inf = meth["__inference"]
    with_node = next(n for n in ast.walk(inf) if isinstance(n, ast.With))
    body = with_node.body
    i0 = next(i for i, s in enumerate(body) if isinstance(s, ast.If) and "prev_d_mean" in ast.unparse(s.test))
    i1 = next(i for i, s in enumerate(body) if isinstance(s, ast.Assign) and ast.unparse(s.targets[0]) == "pred_dur")
    stmts = body[i0:i1 + 1]
    assert stmts[0].lineno == 248 and stmts[-1].lineno == 257, (stmts[0].lineno, stmts[-1].lineno)
    ns = {"torch": torch}
    exec(compile(ast.Module(body=[outl], type_ignores=[]), REF, "exec"), ns)
    code = compile(ast.Module(body=stmts, type_ignores=[]), REF, "exec")
    return ns["__replace_outliers_zscore"], code


def main():
    outlier_fn, code = lift()

    class Holder:
        pass

    holder = Holder()
    setattr(holder, "__replace_outliers_zscore", lambda x: outlier_fn(holder, x))
    rng = np.random.RandomState(7)
    cases = []
    #        L    t    speed prev  spikes
    spec = [(24, 0.1, 1.0, 0.0, 0), (64, 0.1, 1.0, 0.0, 2), (40, 0.1, 1.3, 3.7, 1), (17, 0.0, 0.8, 0.0, 1), (5, 0.1, 1.0, 0.0, 0),
            (4, 0.1, 1.0, 2.5, 0), (3, 0.1, 1.0, 0.0, 0), (2, 0.1, 2.0, 0.0, 0), (200, 0.3, 0.9, 4.1, 3)]
    out = {}
    for k, (L, t, speed, prev, spikes) in enumerate(spec):
        dur = (1.0 + 6.0 * rng.rand(1, L)).astype(np.float32)
        for j in range(spikes):                                   # durations far outside 3 sigma of the rest
            dur[0, 2 + (j * 7) % max(L - 5, 1)] = 40.0 + 5 * j
        ns = {"torch": torch, "self": holder, "duration": torch.from_numpy(dur.copy()), "prev_d_mean": prev, "t": t, "speed": speed,
              "device": "cpu"}
        torch.manual_seed(1000 + k)
        exec(code, ns)                                            # inference.py:248-257, unmodified
        torch.manual_seed(1000 + k)
        z = torch.empty(1, L).normal_(0, 1)
        out["duration_%d" % k] = dur
        out["noise_%d" % k] = z.numpy()
        out["t_%d" % k] = np.float32(t)
        out["speed_%d" % k] = np.float32(speed)
        out["prev_%d" % k] = np.float32(prev)
        out["out_%d" % k] = ns["duration"].numpy().copy()
        out["pred_%d" % k] = np.atleast_1d(ns["pred_dur"].numpy()).astype(np.int32)
        out["mean_%d" % k] = np.float32(ns["duration"].mean())
        cases.append((L, float(np.abs(ns["duration"].numpy() - dur).max())))
    out["n_cases"] = np.int32(len(spec))
    np.savez_compressed(os.path.join(HERE, "smooth_cases.npz"), **out)
    print("smooth", cases)


if __name__ == "__main__":
    main()
