"""Generate diffsims_b200/data/scattering_params.json.

The two electron-scattering-factor parameterisations the hot path uses are
published constants:

* "xtables": International Tables for Crystallography Vol. C, table 4.3.2.3
  (5-Gaussian fits), 98 elements.
* "lobato":  Lobato & Van Dyck, Acta Cryst. A70 (2014) 636-649, 103 elements.

The reference keeps them as Python dict literals
(diffsims/utils/atomic_scattering_params.py:21,
 diffsims/utils/lobato_scattering_params.py:24).  This script loads those two
modules *by file path* in the development container (they import nothing) and
re-serialises the numbers into one compact JSON document: element ->
[a1, b1, ..., a5, b5].  It is run once; the JSON is what ships.

    python tools/make_scattering_tables.py [/root/reference]
"""
import importlib.util
import json
import sys
from pathlib import Path

ref = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
out = Path(__file__).resolve().parents[1] / "diffsims_b200" / "data" / "scattering_params.json"


def load(path, attr):
    spec = importlib.util.spec_from_file_location("_tbl", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return getattr(mod, attr)


xt = load(ref / "diffsims/utils/atomic_scattering_params.py", "ATOMIC_SCATTERING_PARAMS")
lo = load(ref / "diffsims/utils/lobato_scattering_params.py", "ATOMIC_SCATTERING_PARAMS_LOBATO")


def flat(tbl):
    return {el: [float(v) for pair in rows for v in pair] for el, rows in tbl.items()}


doc = {
    "_provenance": {
        "xtables": "International Tables for Crystallography Vol. C, table 4.3.2.3",
        "lobato": "Lobato & Van Dyck, Acta Cryst. A70 (2014) 636-649",
        "layout": "element -> [a1,b1,a2,b2,a3,b3,a4,b4,a5,b5]",
    },
    "xtables": flat(xt),
    "lobato": flat(lo),
}
out.write_text(json.dumps(doc, separators=(",", ":")))
print(out, len(doc["xtables"]), len(doc["lobato"]))
