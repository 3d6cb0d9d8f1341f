"""Golden vectors for the cold-plasma dispersion roots the launchers use (solve_nx_vs_ny_nz_by_bz ->
solve_n1_vs_n2_n3 -> solve_cold_n1sq_vs_n3, RLSDP_cold, slab_eq) from the reference's own output: the slab example
ships vector PDFs of `kx_profiles_slab.<label>` (post_process_lib/slab_processor_m.f90:729-827: k0*nx of the plus/minus
and fast/slow roots at 101 x positions across the slab, written in single precision), one page per ray and root pair.
The pages come from a PDF writer that outlines the tick labels, so the y calibration (2 numbers per page) is not
readable; the test fits those 2 numbers to 404 plotted values per page.  Coordinates are kept in PDF points (1e-4 pt).

    python tests/golden/make_ref_kx_vectors.py   ->  tests/golden/ref_kx_profiles.json
"""
import json
import os
import re
import zlib

REF = "/root/reference/examples_RAYS/ECH_90GHz_slab/pdf_plots"
HERE = os.path.dirname(os.path.abspath(__file__))
# (pdf, page) -> (namelist, ray (1-based), root pair); identified by tests/test_reference_plots.py itself: the
# assignment below is the only one under which a page fits at all (any other misses by tens of points)
PAGES = [("kx_plots.run_1.pdf", 0, "examples/slab_ECH_90GHz_case_1.in", 1, "plus_minus"),
         ("kx_plots.run_1.pdf", 1, "examples/slab_ECH_90GHz_case_1.in", 2, "plus_minus"),
         ("kx_plots.run_1.pdf", 2, "examples/slab_ECH_90GHz_case_1.in", 3, "fast_slow"),
         ("kx_plots.run_2.pdf", 0, "examples/slab_ECH_90GHz_case_2.in", 1, "plus_minus"),
         ("kx_plots.run_2.pdf", 1, "examples/slab_ECH_90GHz_case_2.in", 1, "fast_slow"),
         ("kx_plots.run_2.pdf", 2, "examples/slab_ECH_90GHz_case_2.in", 2, "plus_minus"),
         ("kx_plots.run_2.pdf", 3, "examples/slab_ECH_90GHz_case_2.in", 2, "fast_slow"),
         # run 3: the input is not shipped; reconstructed from these very pages (make_ref_plot_vectors.py), the ray figure is the independent check
         ("kx_plots.run_3.pdf", 0, "examples/slab_ECH_90GHz_case_3.in", 1, "plus_minus"),
         ("kx_plots.run_3.pdf", 1, "examples/slab_ECH_90GHz_case_3.in", 1, "fast_slow"),
         ("kx_plots.run_3.pdf", 2, "examples/slab_ECH_90GHz_case_3.in", 2, "plus_minus"),
         ("kx_plots.run_3.pdf", 3, "examples/slab_ECH_90GHz_case_3.in", 2, "fast_slow"),
         ("kx_plots.run_3.pdf", 4, "examples/slab_ECH_90GHz_case_3.in", 3, "plus_minus"),
         ("kx_plots.run_3.pdf", 5, "examples/slab_ECH_90GHz_case_3.in", 3, "fast_slow")]
# pages whose fast/slow curves carry the labels of the current source (ordered comparison with signs); on the two nz = 0.6
# pages of runs 1 and 2 the complex pair is labelled the other way round (an older build drew them)
UNORDERED = {("kx_plots.run_1.pdf", 2), ("kx_plots.run_2.pdf", 3)}


def page_curves(pdf):
    d = open(pdf, "rb").read()
    out = []
    for s in re.findall(rb"stream\r?\n(.*?)\r?\nendstream", d, re.S):
        try:
            t = zlib.decompress(s).decode("latin1")
        except Exception:
            continue
        paths, cur, nums, col = [], [], [], None
        for tk in t.split():
            try:
                nums.append(float(tk))
                continue
            except ValueError:
                pass
            if tk == "m":
                cur = [tuple(nums[-2:])]
            elif tk in ("l", "c"):
                cur.append(tuple(nums[-2:]))
            elif tk in ("S", "f", "B", "f*", "s"):
                if cur:
                    paths.append((tk, col, cur))
                cur = []
            elif tk in ("SC", "sc", "RG", "rg"):
                col = tuple(nums[-3:])
            nums = []
        big = [(c, p) for op, c, p in paths if op == "S" and len(p) == 101]
        if big:
            out.append(big)
    return out


def main():
    out = {"_doc": "made by tests/golden/make_ref_kx_vectors.py; curve order per page: first root re, im; second root re, im (PDF points)",
           "pages": []}
    cache = {}
    for pdf, ip, nml, ray, pair in PAGES:
        if pdf not in cache:
            cache[pdf] = page_curves(os.path.join(REF, pdf))
        pg = cache[pdf][ip]
        assert len(pg) == 4 and [c for c, _ in pg] == [(1.0, 0.0, 0.0)] * 2 + [(0.0, 0.0, 1.0)] * 2
        out["pages"].append({"pdf": "examples_RAYS/ECH_90GHz_slab/pdf_plots/" + pdf, "page": ip, "namelist": nml, "ray": ray, "roots": pair, "ordered": (pdf, ip) not in UNORDERED,
                             "x_pt": [p[0] for p in pg[0][1]], "curves_pt": [[p[1] for p in c] for _, c in pg]})
    json.dump(out, open(os.path.join(HERE, "ref_kx_profiles.json"), "w"))
    print(len(out["pages"]), "pages")


if __name__ == "__main__":
    main()
