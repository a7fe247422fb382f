"""TEST FIXTURE GENERATOR (runs here, where /root/reference exists): the reference program itself -
oracle/_ref/ben_capture.exe = unmodified main.c with every output() row recorded at full precision - on the four
shipped sample decks and on the generated arc-length shell cap.  Each fixture tests/golden/drv_<name>.npz holds
the deck text (input data, so that the GPU box, which has no /root/reference, can feed it to
oracle/_ref/ben_b200_capture.exe: the same driver on the device path) and the recorded history.

    python tests/golden/make_driver_runs.py
"""
import os
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cu-bens_b200", "python"))
sys.path.insert(0, HERE)
DECKS = "/root/reference/Sample_Input_Files/"


def deck_texts():
    from make_golden import arc_model, ARC, ARC_SHELL
    from cubens_b200 import deck
    out = {}
    for name, fn in (("5a_truss", "model_def_5a_truss.txt"), ("5b_frame", "model_def_5b_frame.txt"),
                     ("5c_shell", "model_def_5c_shell.txt"), ("5d_shell", "model_def_5d_shell.txt")):
        text = open(DECKS + fn).read().replace("\r", "")
        if name == "5d_shell":      # shipped with RFLAG = 1 (restart from a results8.txt that is not shipped)
            lines = text.split("\n"); assert lines[2] == "4,1"; lines[2] = "4,0"; text = "\n".join(lines)
        out[name] = text
    m = arc_model(); c = m.meta["centre"]; a = ARC
    tail = [f"{c},3,{a['dk']!r}", repr(a["alpha"]), repr(a["psi_thresh"]), str(a["iteopt"]),
            f"{a['lpfmax']!r},{a['dkimax']!r}", f"{a['itemax']},{a['submax']},{a['imagmax']},{a['negmax']}",
            f"{a['toldisp']!r},{a['tolforc']!r},{a['tolener']!r}"]
    out["arc_shell"] = deck.write_shell_deck(m, ARC_SHELL, [(c, 3, -1.0e5)], tail)
    return out


def run(exe, text):
    with tempfile.TemporaryDirectory() as td:
        open(os.path.join(td, "model_def.txt"), "w").write(text)
        subprocess.run([exe], cwd=td, stdout=subprocess.DEVNULL, timeout=300, check=True)
        raw = open(os.path.join(td, "capture.bin"), "rb").read()
        ok = "Solution successful" in open(os.path.join(td, "results1.txt")).read()
    neq, nrows = np.frombuffer(raw[:16], dtype=np.int64)
    return np.frombuffer(raw[16:], dtype=np.float64).reshape(nrows, neq + 2).copy(), ok


if __name__ == "__main__":
    exe = os.path.join(ROOT, "oracle", "_ref", "ben_capture.exe")
    for name, text in deck_texts().items():
        hist, ok = run(exe, text)
        np.savez_compressed(os.path.join(HERE, f"drv_{name}.npz"), deck_text=np.frombuffer(text.encode(), dtype=np.uint8),
                            hist=hist, ok=np.array(ok))
        print(name, "rows", hist.shape, "successful", ok)
