"""Records what the reference tree itself pins of the MX path (run in the build container, where /root/reference exists):
the element-format parameters of formats.py (_get_format_params, ElemFormat values) and finalize_mx_specs(apply_mx_specs(.)) of
specs.py on the configurations bfp_util.extract_mx_args produces.  Output: tests/golden/mx_reference_tables.json.
    python tests/golden/make_mx_golden.py"""
import importlib.util
import json
import os

REF = "/root/reference/src/transformers/bfp"
HERE = os.path.dirname(os.path.abspath(__file__))


def load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    fm, sp = load("formats"), load("specs")
    formats = {}
    for name in ("int8", "int4", "int2", "fp8_e5m2", "fp8_e4m3", "fp6_e3m2", "fp6_e2m3", "fp4_e2m1", "fp4"):
        ebits, mbits, emax, max_norm, min_norm = fm._get_format_params(name)
        formats[name] = {"id": fm.ElemFormat.from_str(name).value, "ebits": ebits, "mbits": mbits, "emax": emax, "max_norm": float(max_norm),
                         "min_norm": float(min_norm)}
    specs = []
    for given in (None, {}, {"bfloat": 16},
                  {"w_elem_format": "fp8_e4m3", "a_elem_format": "fp8_e4m3", "block_size": 32, "bfloat": 16, "scale_bits": 8},
                  {"w_elem_format": "fp4_e2m1", "a_elem_format": "fp6_e2m3", "block_size": 64, "bfloat": 16, "scale_bits": 8},
                  {"a_elem_format": "int8", "scale_bits": 8, "block_size": 64, "round": "floor"}):
        out = sp.finalize_mx_specs(sp.apply_mx_specs(dict(given) if given is not None else None))
        specs.append({"given": given, "finalized": dict(out) if out is not None else None})
    with open(os.path.join(HERE, "mx_reference_tables.json"), "w") as f:
        json.dump({"source": "formats.py:52-128, specs.py:172-279 of the reference tree", "formats": formats, "specs": specs}, f, indent=1, sort_keys=True)
    print("wrote", len(formats), "formats,", len(specs), "spec cases")


if __name__ == "__main__":
    main()
