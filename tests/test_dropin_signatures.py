"""The drop-in headers must declare every public (and protected) member of the reference's ORBextractor / ORBmatcher with the SAME
signature (R/include/ORBextractor.h:47-113, R/include/ORBmatcher.h:35-108): the reference's callers are recompiled against them
unchanged.  Parses both header pairs; needs /root/reference (skipped on the GPU box)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFINC = "/root/reference/src/orb_slam3_ros/orb_slam3/include"

pytestmark = pytest.mark.skipif(not os.path.exists(os.path.join(REFINC, "ORBmatcher.h")), reason="/root/reference is absent")


def class_members(path, cls):
    """{access: [normalised member declarations]} of class `cls` (comments stripped, inline bodies dropped)."""
    txt = open(path, encoding="utf-8", errors="replace").read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    txt = re.sub(r"//[^\n]*", " ", txt)
    m = re.search(r"\bclass\s+%s\b[^;{]*\{" % cls, txt)
    assert m, (path, cls)
    depth, i = 1, m.end()
    while depth:
        depth += {"{": 1, "}": -1}.get(txt[i], 0)
        i += 1
    body = txt[m.end():i - 1]
    # drop inline function bodies, keep the signature
    out, depth, cur = [], 0, ""
    for ch in body:
        if ch == "{":
            depth += 1
            if depth == 1:
                cur += ";"
            continue
        if ch == "}":
            depth -= 1
            continue
        if depth == 0:
            cur += ch
    access, members = "private", {"public": [], "protected": [], "private": []}
    for stmt in re.split(r";", cur):
        stmt = stmt.strip()
        while True:
            a = re.match(r"(public|protected|private)\s*:\s*", stmt)
            if not a:
                break
            access = a.group(1)
            stmt = stmt[a.end():]
        stmt = re.sub(r"\s+", " ", stmt).strip()
        stmt = re.sub(r"\s*([(),&*<>=])\s*", r"\1", stmt)
        stmt = re.sub(r"\bstd::", "", stmt)                 # the reference mixes std::vector and vector (using namespace std)
        stmt = re.sub(r"\binline\b ?", "", stmt).replace("int GetLevels", "int GetLevels")
        if stmt:
            members[access].append(stmt)
    return members


@pytest.mark.parametrize("header,cls,min_public", [("ORBmatcher.h", "ORBmatcher", 18), ("ORBextractor.h", "ORBextractor", 10)])
def test_every_reference_member_is_declared_identically(header, cls, min_public):
    ref = class_members(os.path.join(REFINC, header), cls)
    mine = class_members(os.path.join(ROOT, "dropin", header), cls)
    assert len(ref["public"]) >= min_public, ref["public"]
    for access in ("public", "protected"):
        for decl in ref[access]:
            if cls == "ORBextractor" and access == "protected":
                continue              # the private part is the C-ABI handle instead of the reference's tables (documented in the header)
            if decl.startswith("~ORBextractor"):
                assert any(d.startswith("~ORBextractor") for d in mine[access]), decl
                continue
            assert decl in mine[access], "%s member missing or different in dropin/%s:\n  %s\nhave:\n  %s" % (
                access, header, decl, "\n  ".join(mine[access]))


def test_matcher_has_all_fifteen_public_methods():
    mine = class_members(os.path.join(ROOT, "dropin", "ORBmatcher.h"), "ORBmatcher")
    names = [re.match(r".*?\b(\w+)\(", d).group(1) for d in mine["public"] if "(" in d]
    want = {"ORBmatcher": 1, "DescriptorDistance": 1, "SearchByProjection": 5, "SearchByBoW": 2, "SearchForInitialization": 1,
            "SearchForTriangulation": 2, "SearchBySim3": 1, "Fuse": 2}
    for n, c in want.items():
        assert names.count(n) == c, (n, names.count(n))
    prot = [re.match(r".*?\b(\w+)\(", d).group(1) for d in mine["protected"] if "(" in d]
    assert sorted(prot) == ["CheckDistEpipolarLine", "CheckDistEpipolarLine2", "ComputeThreeMaxima", "RadiusByViewingCos"]
