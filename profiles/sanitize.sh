# memcheck of the history kernels on small cases (QA, not a benchmark)
cat > /tmp/san.py <<'PY'
import os, sys
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import scone_b200
R = os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
for deck, ov, fixed in ((R + "/decks/ce/pincell", "pop 1500; inactive 1; active 1; seed 3;", False),
                        (R + "/decks/ce/assembly17", "pop 1500; inactive 1; active 1; seed 3;", False),
                        (R + "/decks/c5g7/c5g7_2d", "pop 3000; inactive 1; active 1; seed 3; transportOperator { type transportOperatorST; }", False),
                        (R + "/decks/c5g7/c5g7_2d", "pop 3000; inactive 1; active 1; seed 3;", False),
                        (R + "/decks/fixed/ce_sphere", "pop 2000; cycles 2; seed 3;", True),
                        (R + "/decks/fixed/mg_sphere", "pop 2000; cycles 2; seed 3;", True),
                        (R + "/decks/fixed/mg_sphere", "pop 2000; cycles 2; seed 3; source { type materialSource; mat UO2; data mg; G 2; }", True),
                        (R + "/decks/mg/can", "pop 3000; inactive 1; active 1; seed 3;", False),
                        (R + "/decks/mg/can", "pop 3000; inactive 1; active 1; seed 3; transportOperator { type transportOperatorDT; }", False),
                        (R + "/decks/c5g7/c5g7_2d", "pop 3000; inactive 1; active 1; seed 3; printSource 2; outputFile /tmp/san_dump; "
                         "activeTally { e { type shannonEntropyClerk; cycles 1; map { type spaceMap; axis x; grid lin; min -32.13; max 32.13; N 10; } } }", False)):
    for cfg in (("", "events") if "ce/" in deck and not fixed else ("",)):
        if cfg: os.environ["SB_CE_KERNEL"] = cfg
        else: os.environ.pop("SB_CE_KERNEL", None)
        pp = scone_b200.EigenPhysicsPackage(deck, ov, device=0)
        if fixed:
            for _ in range(2): pp.fixed_cycle()
        else:
            pp.generateInitialState(); pp.cycle(False); pp.cycle(True)
        print("ok", deck.split("/")[-1], cfg or "default", pp.k)
        pp.close()
PY
compute-sanitizer --tool ${SAN_TOOL:-memcheck} --print-limit 5 python /tmp/san.py 2>&1 | tail -25
