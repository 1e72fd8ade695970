# memcheck / racecheck / synccheck of the delta-tracking kernels (k_histories + k_lone) on small cases (QA, not a benchmark)
cat > /tmp/san_dt.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200
R = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for deck, ov in ((R + "/decks/c5g7/c5g7_2d", "pop 3000; inactive 1; active 1; seed 3;"), (R + "/decks/c5g7/c5g7_3d_rodded", "pop 2000; inactive 1; active 1; seed 3;"),
                 (R + "/decks/urr/inf", "pop 2000; inactive 1; active 1; seed 3;"), (R + "/decks/mg/can", "pop 2000; inactive 1; active 1; seed 3; transportOperator { type transportOperatorDT; }")):
    for t in ("", "1", "32"):
        if t: os.environ["SB_ASSIST"] = t
        else: os.environ.pop("SB_ASSIST", None)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        pp.generateInitialState(); pp.cycle(False); pp.cycle(True)
        print("ok", deck.split("/")[-1], "SB_ASSIST=" + (t or "auto"), pp.k)
        pp.close()
PY
for tool in ${SAN_TOOLS:-memcheck racecheck synccheck}; do echo "== $tool"; compute-sanitizer --tool $tool --print-limit 5 python /tmp/san_dt.py 2>&1 | grep -v "^ok" | tail -8; done
