"""Write the overlay headers of SURVEY.md section 8b option 1: operators/softmax.h and operators/hist.h of the reference with
their two one-line changes (the transform node hands its impl the EXECUTOR instead of the bare stream, so an executor type
can be told apart).  The reference's text is read where it lies (MATX_REFERENCE, default /root/reference), patched in
memory and written under oracle/_ref/overlay/ — a build artefact like the rest of oracle/_ref (git-ignored): nothing of the
reference is committed here.  An include directory `-I oracle/_ref/overlay` in front of the reference's makes
`(out = softmax(x)).run(exec)` and `(out = hist(x, lo, hi, levels)).run(exec)` reach include/matx_b200/executor.h.

    python tools/make_overlay.py            # prints the directory
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("MATX_REFERENCE", "/root/reference")
OUT = os.path.join(ROOT, "oracle", "_ref", "overlay")

PATCHES = {
    "matx/operators/softmax.h": [("softmax_impl(cuda::std::get<0>(out), a_, perm_, ex.getStream());", "softmax_impl(cuda::std::get<0>(out), a_, perm_, ex);"),
                                 ("softmax_impl(cuda::std::get<0>(out), a_, ex.getStream());", "softmax_impl(cuda::std::get<0>(out), a_, ex);")],
    "matx/operators/hist.h": [("hist_impl(cuda::std::get<0>(out), a_, lower_, upper_, num_levels_, ex.getStream());",
                               "hist_impl(cuda::std::get<0>(out), a_, lower_, upper_, num_levels_, ex);")],
}


def main() -> str:
    for rel, subs in PATCHES.items():
        src = os.path.join(REF, "include", rel)
        text = open(src).read()
        for old, new in subs:
            if text.count(old) != 1:
                raise SystemExit("%s: expected exactly one occurrence of %r (the reference changed: update tools/make_overlay.py)" % (rel, old))
            text = text.replace(old, new)
        dst = os.path.join(OUT, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        with open(dst, "w") as f:
            f.write(text)
    return OUT


if __name__ == "__main__":
    print(main())
    sys.exit(0)
