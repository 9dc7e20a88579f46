#!/usr/bin/env python3
"""Write oracle/_ref/kernel_hooked.cpp: the reference's source/kernel.cpp with the inline narrowphase of dispatch()
replaced by a call to mcb200_hook_narrowphase() (mcut_b200/csrc/shim/mcut_hook.h).

The reference source is read where it lies (/root/reference), the OUTPUT goes to the git-ignored oracle/_ref/ only; the
text that is inserted is this repo's own (mcut_b200/csrc/shim/kernel_hook_region.inc).  The region is found by its
markers, not by line numbers:
    start: the banner comment "Calculate polygon intersection points" above the declaration of
           ps_edge_face_intersection_pairs (kernel.cpp:1775-1779)
    end  : the comment "Create edges from the new intersection points" (kernel.cpp:3208)
usage: make_hooked_kernel.py <reference root> <region.inc> <out.cpp>"""
import sys

ref, inc, out = sys.argv[1], sys.argv[2], sys.argv[3]
src = open(ref + "/source/kernel.cpp", encoding="utf-8", errors="replace").read().split("\n")
decl = [i for i, l in enumerate(src) if "ps_edge_face_intersection_pairs;" in l and "std::unordered_map<ed_t" in l and not l.strip().startswith("//")]
assert len(decl) == 1, "start marker not found exactly once: %r" % decl
start = decl[0]
while start > 0 and "Calculate polygon intersection points" not in src[start]:
    start -= 1
assert start > 0
start -= 1  # the opening line of the banner
end = [i for i, l in enumerate(src) if l.strip() == "// Create edges from the new intersection points"]
assert len(end) == 1 and end[0] > decl[0], "end marker not found exactly once: %r" % end
region = open(inc, encoding="utf-8").read().rstrip("\n").split("\n")
marker = 'extern "C" { int mcb200_kernel_is_hooked = 1; } // tells the adapter that nobody will read face_bboxes on the host'
text = ['#include "mcut_hook.h" // mcut_b200 narrowphase hook', marker] + src[:start] + region + src[end[0]:]
open(out, "w", encoding="utf-8").write("\n".join(text))
print("kernel_hooked.cpp: replaced reference lines %d..%d (%d lines) by %d lines" % (start + 1, end[0], end[0] - start, len(region)))
