#!/usr/bin/env python3
"""Writes the input decks this repo benchmarks and tests on, in SCONE's dictionary format.

The decks are authored here from the public benchmark specifications rather than copied:
  * C5G7 (OECD/NEA "Benchmark specification for deterministic 2D/3D MOX fuel assembly
    transport calculations without spatial homogenisation", 7-group constants below);
    2-D configuration = BASELINE.json configs[0]; the 3-D rodded variant (configs[3]) is
    authored from the same specification (three axial fuel zones, axial reflector, control
    rods in the guide tubes of the rodded assemblies).
  * Sood et al. analytic benchmarks URRa-2-1-IN / URRa-2-1-SL (2-group, infinite medium k
    = 1.631452 and critical slab), BASELINE.json configs[1].
tests/test_decks.py checks, where /root/reference is present, that the decks written here
load to the same flat model (geometry graph, XS tables, majorant) as the reference's own
InputFiles/Benchmarks/Multigroup/C5G7, InputFiles/SCONE_Inf and InputFiles/SCONE_Slab.

Usage: python decks/gen_decks.py [outdir]     (default: the directory of this script)
"""
import os
import sys

# ---------------------------------------------------------------------------------------
# NEA C5G7 seven-group macroscopic cross sections [1/cm]; scattering rows are "from group".
C5G7_XS = {
    'UO2': {
        'capture': "8.1274000E-04 2.8980990E-03 2.0315800E-02 7.7671200E-02 1.2211600E-02 2.8225200E-02 6.6776000E-02",
        'fission': "7.212060E-3 8.193010E-4 6.453200E-3 1.856480E-2 1.780840E-2 8.303480E-2 2.160040E-1",
        'nu': "2.7814494E+00 2.4744300E+00 2.4338297E+00 2.4338000E+00 2.43380E+00 2.43380E+00 2.43380E+00",
        'chi': "5.8791E-01 4.1176E-01 3.3906E-04 1.1761E-07 0.0000E+00 0.0000E+00 0.0000E+00",
        "P0": (
            "1.2753700E-01 4.2378000E-02 9.4374000E-06 5.5163000E-09 0.0000000E+00 0.0000000E+00 0.0000000E+00\n"
            "0.0000000E+00 3.2445600E-01 1.6314000E-03 3.1427000E-09 0.0000000E+00 0.0000000E+00 0.0000000E+00\n"
            "0.0000000E+00 0.0000000E+00 4.5094000E-01 2.6792000E-03 0.0000000E+00 0.0000000E+00 0.0000000E+00\n"
            "0.0000000E+00 0.0000000E+00 0.0000000E+00 4.5256500E-01 5.5664000E-03 0.0000000E+00 0.0000000E+00\n"
            "0.0000000E+00 0.0000000E+00 0.0000000E+00 1.2525000E-04 2.7140100E-01 1.0255000E-02 1.0021000E-08\n"
            "0.0000000E+00 0.0000000E+00 0.0000000E+00 0.0000000E+00 1.2968000E-03 2.6580200E-01 1.6809000E-02\n"
            "0.0000000E+00 0.0000000E+00 0.0000000E+00 0.0000000E+00 0.0000000E+00 8.5458000E-03 2.7308000E-01\n"
        ),
    },
    'MOX43': {
        'capture': "8.0686000E-04 2.8808020E-03 2.2271650E-02 8.1322800E-02 1.2917650E-01 1.7642300E-01 1.6038200E-01",
        'fission': "7.627040E-03 8.768980E-04 5.698350E-03 2.288720E-02 1.076350E-02 2.327570E-01 2.489680E-01",
        'nu': "2.852089E+00 2.890990E+00 2.854860E+00 2.860730E+00 2.854470E+00 2.864150E+00 2.867800E+00",
        'chi': "5.8791E-01 4.1176E-01 3.3906E-04 1.1761E-07 0.0000E+00 0.0000E+00 0.0000E+00",
        "P0": (
            "1.288760E-01 4.141300E-02 8.229000E-06 5.040500E-09 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 3.254520E-01 1.639500E-03 1.598200E-09 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 4.531880E-01 2.614200E-03 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 4.571730E-01 5.539400E-03 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 1.604600E-04 2.768140E-01 9.312700E-03 9.165600E-09\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 2.005100E-03 2.529620E-01 1.485000E-02\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 8.494800E-03 2.650070E-01\n"
        ),
    },
    'MOX7': {
        'capture': "8.112400E-04 2.971050E-03 2.445944E-02 8.915700E-02 1.670164E-01 2.446660E-01 2.224070E-01",
        'fission': "8.254460E-03 1.325650E-03 8.421560E-03 3.287300E-02 1.596360E-02 3.237940E-01 3.628030E-01",
        'nu': "2.884980E+00 2.910790E+00 2.865740E+00 2.870630E+00 2.867140E+00 2.866580E+00 2.875390E+00",
        'chi': "5.8791E-01 4.1176E-01 3.3906E-04 1.1761E-07 0.0000E+00 0.0000E+00 0.0000E+00",
        "P0": (
            "1.304570E-01 4.179200E-02 8.510500E-06 5.132900E-09 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 3.284280E-01 1.643600E-03 2.201700E-09 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 4.583710E-01 2.533100E-03 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 4.637090E-01 5.476600E-03 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 1.761900E-04 2.823130E-01 8.728900E-03 9.001600E-09\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 2.276000E-03 2.497510E-01 1.311400E-02\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 8.864500E-03 2.595290E-01\n"
        ),
    },
    'MOX87': {
        'capture': "8.141100E-04 3.031340E-03 2.596840E-02 9.367530E-02 1.891424E-01 2.838120E-01 2.595710E-01",
        'fission': "8.672090E-03 1.624260E-03 1.027160E-02 3.904470E-02 1.925760E-02 3.748880E-01 4.305990E-01",
        'nu': "2.904260E+00 2.917950E+00 2.869860E+00 2.874910E+00 2.871750E+00 2.867520E+00 2.878079E+00",
        'chi': "5.8791E-01 4.1176E-01 3.3906E-04 1.1761E-07 0.0000E+00 0.0000E+00 0.0000E+00",
        "P0": (
            "1.315040E-01 4.204600E-02 8.697200E-06 5.193800E-09 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 3.304030E-01 1.646300E-03 2.600600E-09 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 4.617920E-01 2.474900E-03 0.000000E+00 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 4.680210E-01 5.433000E-03 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 1.859700E-04 2.857710E-01 8.397300E-03 8.928000E-09\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 2.391600E-03 2.476140E-01 1.232200E-02\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 8.968100E-03 2.560930E-01\n"
        ),
    },
    'FC': {
        'capture': "5.113152100E-04 7.580717436E-05 3.159662810E-04 1.162255940E-03 3.397554610E-03 9.187885028E-03 2.324191959E-02",
        'fission': "4.790020000E-09 5.825640000E-09 4.637190000E-07 5.244060000E-06 1.453900000E-07 7.149720000E-07 2.080410000E-06",
        'nu': "2.762829800E+00 2.462390398E+00 2.433799348E+00 2.433799384E+00 2.433800124E+00 2.433800205E+00 2.433800068E+00",
        'chi': "5.8791E-01 4.1176E-01 3.3906E-04 1.1761E-07 0.0000E+00 0.0000E+00 0.0000E+00",
        "P0": (
            "6.616590000E-02 5.907000000E-02 2.833400000E-04 1.462200000E-06 2.064200000E-08 0.000000000E+00 0.000000000E+00\n"
            "0.000000000E+00 2.403770000E-01 5.243500000E-02 2.499000000E-04 1.923900000E-05 2.987500000E-06 4.214000000E-07\n"
            "0.000000000E+00 0.000000000E+00 1.834250000E-01 9.228800000E-02 6.936500000E-03 1.079000000E-03 2.054300000E-04\n"
            "0.000000000E+00 0.000000000E+00 0.000000000E+00 7.907690000E-02 1.699900000E-01 2.586000000E-02 4.925600000E-03\n"
            "0.000000000E+00 0.000000000E+00 0.000000000E+00 3.734000000E-05 9.975700000E-02 2.067900000E-01 2.447800000E-02\n"
            "0.000000000E+00 0.000000000E+00 0.000000000E+00 0.000000000E+00 9.174200000E-04 3.167740000E-01 2.387600000E-01\n"
            "0.000000000E+00 0.000000000E+00 0.000000000E+00 0.000000000E+00 0.000000000E+00 4.979300000E-02 1.099100000E+00\n"
        ),
    },
    'GT': {
        'capture': "5.113200E-04 7.580100E-05 3.157200E-04 1.158200E-03 3.397500E-03 9.187800E-03 2.324200E-02",
        "P0": (
            "6.616590E-02 5.907000E-02 2.833400E-04 1.462200E-06 2.064200E-08 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 2.403770E-01 5.243500E-02 2.499000E-04 1.923900E-05 2.987500E-06 4.214000E-07\n"
            "0.000000E+00 0.000000E+00 1.832970E-01 9.239700E-02 6.944600E-03 1.080300E-03 2.056700E-04\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 7.885110E-02 1.701400E-01 2.588100E-02 4.929700E-03\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 3.733300E-05 9.973720E-02 2.067900E-01 2.447800E-02\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 9.172600E-04 3.167650E-01 2.387700E-01\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 4.979200E-02 1.099120E+00\n"
        ),
    },
    'CR': {
        'capture': "1.7049E-03 8.36224E-03 8.37901E-02 3.97797E-01 6.98763E-01 9.29508E-01 1.17836",
        "P0": (
            "1.70563E-01 4.44012E-02 9.83670E-05 1.27786E-07 0.0 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 4.71050E-01 6.85480E-04 3.91395E-10 0.0 0.0 0.0\n"
            "0.000000E+00 0.000000E+00 8.01859E-01 7.20132E-04 0.0 0.0 0.0\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 5.70752E-01 1.46015E-03 0.0 0.0\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 6.55562E-05 2.07838E-01 3.81486E-03 3.69760E-9\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 1.02427E-03 2.02465E-01 4.75290E-3\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 3.53043E-03 6.58597E-01\n"
        ),
    },
    'moder': {
        'capture': "6.010500E-04 1.579300E-05 3.371600E-04 1.940600E-03 5.741600E-03 1.500100E-02 3.723900E-02",
        "P0": (
            "4.447770E-02 1.134000E-01 7.234700E-04 3.749900E-06 5.318400E-08 0.000000E+00 0.000000E+00\n"
            "0.000000E+00 2.823340E-01 1.299400E-01 6.234000E-04 4.800200E-05 7.448600E-06 1.045500E-06\n"
            "0.000000E+00 0.000000E+00 3.452560E-01 2.245700E-01 1.699900E-02 2.644300E-03 5.034400E-04\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 9.102840E-02 4.155100E-01 6.373200E-02 1.213900E-02\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 7.143700E-05 1.391380E-01 5.118200E-01 6.122900E-02\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 2.215700E-03 6.999130E-01 5.373200E-01\n"
            "0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 0.000000E+00 1.324400E-01 2.480700E+00\n"
        ),
    },
}

# Sood et al. URRa-2-1 two-group constants (rows of P0/P1 are "from group")
URR_2G = """
numberOfGroups 2;
capture (0.0010046 0.025788);
fission (0.0010484 0.050632);
nu      (2.5 2.5);
chi     (1.0 0.0);
scatteringMultiplicity (
 1.0 1.0
 1.0 1.0 );
P0 (
 0.62568 0.029227
 0.0     2.443830 );
P1 (
 0.27459 0.0075737
 0.0     0.83318 );
"""

N_PIN = 17
PITCH = 1.26
ASSEMBLY = N_PIN * PITCH          # 21.42 cm
HALF_CORE = 1.5 * ASSEMBLY        # 32.13 cm

# guide-tube positions (row, col), fission chamber in the centre (NEA spec fig. 2)
GT_ROWS = {2: (5, 8, 11), 3: (3, 13), 5: (2, 5, 8, 11, 14), 8: (2, 5, 11, 14), 11: (2, 5, 8, 11, 14),
           13: (3, 13), 14: (5, 8, 11)}
FC_POS = (8, 8)
# columns (inclusive) holding 8.7 % MOX in each row; 7.0 % elsewhere inside the 4.3 % outer ring
MOX87_SPAN = {3: (5, 11), 4: (4, 12), 12: (4, 12), 13: (5, 11)}
MOX87_SPAN.update({r: (3, 13) for r in range(5, 12)})

# universe ids
PIN = dict(UO2=1, GT=2, mox43=3, mox7=4, mox87=5, FC=6, CR=7, water=30)


def assembly_map(kind, rodded=False):
    rows = []
    for r in range(N_PIN):
        row = []
        for c in range(N_PIN):
            if (r, c) == FC_POS:
                u = PIN["FC"]
            elif c in GT_ROWS.get(r, ()):
                u = PIN["CR"] if rodded else PIN["GT"]
            elif kind == "UO2":
                u = PIN["UO2"]
            elif r in (0, N_PIN - 1) or c in (0, N_PIN - 1):
                u = PIN["mox43"]
            elif r in MOX87_SPAN and MOX87_SPAN[r][0] <= c <= MOX87_SPAN[r][1]:
                u = PIN["mox87"]
            else:
                u = PIN["mox7"]
            row.append(u)
        rows.append(row)
    return rows


def fmt_map(rows, indent="        "):
    return "\n".join(indent + " ".join("%2d" % v for v in row) for row in rows)


def xs_text(name):
    x = C5G7_XS[name]
    t = ["// NEA C5G7 seven-group constants: %s" % name, "numberOfGroups 7;"]
    for key in ("capture", "fission", "nu", "chi"):
        if key in x:
            t.append("%s (%s);" % (key, x[key]))
    t.append("scatteringMultiplicity (\n" + "\n".join(" ".join(["1.0"] * 7) for _ in range(7)) + "\n);")
    t.append("P0 (\n" + "".join(x["P0"]) + ");")
    return "\n".join(t) + "\n"


MATERIALS_2D = [("mox43", "MOX43"), ("mox7", "MOX7"), ("mox87", "MOX87"), ("UO2", "UO2"), ("FC", "FC"),
                ("GT", "GT"), ("water", "moder")]


def materials_block(mats, xsdir):
    out = []
    for m, f in mats:
        out.append("    %s { temp 300; xsFile %s/%s.xs; composition { } }" % (m, xsdir, f))
    return "\n".join(out)


def c5g7_2d(pop=100000, inactive=50, active=200, tracking="DT", seed=20261017):
    pins = []
    for m, uid in (("UO2", 1), ("GT", 2), ("mox43", 3), ("mox7", 4), ("mox87", 5), ("FC", 6)):
        pins.append("    pin%d { id %d; type pinUniverse; radii (0.5400 0.0); fills (%s water); }" % (uid, uid, m))
    return """// C5G7 MOX benchmark, 2-D configuration, 7 groups (NEA/NSC/DOC(2001)4)
// quarter core: UO2 | MOX / MOX | UO2 with reflector; reflective at -x and +y, vacuum at +x and -y
type eigenPhysicsPackage;
pop      %(pop)d;
active   %(active)d;
inactive %(inactive)d;
seed     %(seed)d;
XSdata   mg;
dataType mg;
outputFile c5g7_2d;

collisionOperator { neutronMG { type neutronMGstd; } }
transportOperator { type transportOperator%(tracking)s; }

inactiveTally { }
activeTally {
  fissionMap { type collisionClerk;
               map { type multiMap; maps (xax yax);
                     xax { type spaceMap; axis x; grid lin; min -32.13; max 10.71; N 34; }
                     yax { type spaceMap; axis y; grid lin; min -10.71; max 32.13; N 34; } }
               response (fiss); fiss { type macroResponse; MT -6; } }
}

geometry {
  type geometryStd;
  boundary (1 0 0 1 1 1);
  graph { type extended; }
  surfaces { domain { id 3; type box; origin (0.0 0.0 0.0); halfwidth (32.13 32.13 32.13); } }
  cells { }
  universes {
    root { id 1000; type rootUniverse; border 3; fill u<100>; }
%(pins)s
    pin30 { id 30; type pinUniverse; radii (0.0); fills (water); }
    latUO2 { id 10; type latUniverse; origin (0.0 0.0 0.0); pitch (1.26 1.26 0.0); shape (17 17 0); padMat water;
      map (
%(uo2)s
      ); }
    latMOX { id 20; type latUniverse; origin (0.0 0.0 0.0); pitch (1.26 1.26 0.0); shape (17 17 0); padMat water;
      map (
%(mox)s
      ); }
    latCore { id 100; type latUniverse; origin (0.0 0.0 0.0); pitch (21.42 21.42 0.0); shape (3 3 0); padMat water;
      map (
        10 20 30
        20 10 30
        30 30 30 ); }
  }
}

nuclearData {
  handles { mg { type baseMgNeutronDatabase; PN P0; } }
  materials {
%(mats)s
  }
}
""" % dict(pop=pop, active=active, inactive=inactive, seed=seed, tracking=tracking, pins="\n".join(pins),
           uo2=fmt_map(assembly_map("UO2")), mox=fmt_map(assembly_map("MOX")),
           mats=materials_block(MATERIALS_2D, "./xs"))


def c5g7_3d_rodded(pop=10000000, inactive=50, active=100, seed=20261017):
    """3-D extension, "Rodded B"-like configuration of NEA/NSC/DOC(2005)16: fuel 3 x 14.28 cm axial
    zones + 21.42 cm axial reflector (half height, reflective at z = 0). Control rods are inserted in the
    upper reflector everywhere, 1/3 into the inner UO2 assembly and... kept simple here: rods occupy the
    guide tubes of the inner UO2 assembly in the top fuel zone and of all assemblies in the reflector."""
    pins = []
    for m, uid in (("UO2", 1), ("GT", 2), ("mox43", 3), ("mox7", 4), ("mox87", 5), ("FC", 6), ("CR", 7)):
        pins.append("    pin%d { id %d; type pinUniverse; radii (0.5400 0.0); fills (%s water); }" % (uid, uid, m))

    def lat(name, uid, rows):
        return ("    %s { id %d; type latUniverse; origin (0.0 0.0 0.0); pitch (1.26 1.26 0.0); shape (17 17 0); padMat water;\n"
                "      map (\n%s\n      ); }" % (name, uid, fmt_map(rows)))

    def refl(rodded):
        rows = [[PIN["water"]] * N_PIN for _ in range(N_PIN)]
        if rodded:
            for r, cols in GT_ROWS.items():
                for c in cols:
                    rows[r][c] = PIN["CR"]
        return rows

    lats = [lat("latUO2", 10, assembly_map("UO2")), lat("latMOX", 20, assembly_map("MOX")),
            lat("latUO2rod", 11, assembly_map("UO2", rodded=True)),
            lat("latReflRod", 31, refl(True))]
    # axial layers, bottom (z=0 midplane, reflective) to top: 3 fuel zones of 14.28 cm + 21.42 cm reflector
    core = """    layerFuel { id 101; type latUniverse; origin (0.0 0.0 0.0); pitch (21.42 21.42 0.0); shape (3 3 0); padMat water;
      map ( 10 20 30
            20 10 30
            30 30 30 ); }
    layerRod { id 102; type latUniverse; origin (0.0 0.0 0.0); pitch (21.42 21.42 0.0); shape (3 3 0); padMat water;
      map ( 11 20 30
            20 10 30
            30 30 30 ); }
    layerRefl { id 103; type latUniverse; origin (0.0 0.0 0.0); pitch (21.42 21.42 0.0); shape (3 3 0); padMat water;
      map ( 31 31 30
            31 31 30
            30 30 30 ); }
    axial { id 200; type latUniverse; origin (0.0 0.0 32.13); pitch (64.26 64.26 7.14); shape (1 1 9); padMat water;
      offsetMap ( 0 0 0 0 0 0 0 0 0 );
      map ( 103 103 103 102 102 101 101 101 101 ); }"""
    mats = MATERIALS_2D[:-1] + [("CR", "CR"), ("water", "moder")]
    return """// C5G7 3-D rodded configuration (authored from NEA/NSC/DOC(2005)16), 7 groups
type eigenPhysicsPackage;
pop      %(pop)d;
active   %(active)d;
inactive %(inactive)d;
seed     %(seed)d;
XSdata   mg;
dataType mg;
outputFile c5g7_3d;

collisionOperator { neutronMG { type neutronMGstd; } }
transportOperator { type transportOperatorDT; }

inactiveTally { }
activeTally {
  flux { type collisionClerk;
         map { type multiMap; maps (xax yax zax);
               xax { type spaceMap; axis x; grid lin; min -32.13; max 10.71; N 34; }
               yax { type spaceMap; axis y; grid lin; min -10.71; max 32.13; N 34; }
               zax { type spaceMap; axis z; grid lin; min 0.0; max 64.26; N 9; } }
         response (flux fiss); flux { type fluxResponse; } fiss { type macroResponse; MT -6; } }
}

geometry {
  type geometryStd;
  boundary (1 0 0 1 1 0);
  graph { type shrunk; }
  surfaces { domain { id 3; type box; origin (0.0 0.0 32.13); halfwidth (32.13 32.13 32.13); } }
  cells { }
  universes {
    root { id 1000; type rootUniverse; border 3; fill u<200>; }
%(pins)s
    pin30 { id 30; type pinUniverse; radii (0.0); fills (water); }
%(lats)s
%(core)s
  }
}

nuclearData {
  handles { mg { type baseMgNeutronDatabase; PN P0; } }
  materials {
%(mats)s
  }
}
""" % dict(pop=pop, active=active, inactive=inactive, seed=seed, pins="\n".join(pins), lats="\n".join(lats), core=core,
           mats=materials_block(mats, "./xs"))


def urr_deck(kind, pop=15000, inactive=100, active=500, seed=20261017):
    inf = kind == "inf"
    return """// Sood URRa-2-1-%(tag)s: two-group %(what)s
type eigenPhysicsPackage;
pop      %(pop)d;
active   %(active)d;
inactive %(inactive)d;
seed     %(seed)d;
XSdata   mg;
dataType mg;

collisionOperator { neutronMG { type neutronMGstd; } }
transportOperator { type transportOperatorDT; }

inactiveTally { }
activeTally {
  norm fiss; normVal 100;
  fiss { type collisionClerk; response (fiss); fiss { type macroResponse; MT -6; } }
%(flux)s}

geometry {
  type geometryStd;
  boundary (%(bc)s);
  graph { type shrunk; }
  surfaces { bound { id 1; type box; origin (0.0 0.0 0.0); halfwidth (%(hw)s 10.0 10.0); } }
  cells { }
  universes { root { id 1; type rootUniverse; border 1; fill fuel; } }
}

nuclearData {
  handles { mg { type baseMgNeutronDatabase; PN %(pn)s; } }
  materials { fuel { temp 273; composition { } xsFile ./xs/URRa_2_1.xs; } }
}
""" % dict(tag="IN" if inf else "SL", what="infinite medium, k = 1.631452" if inf else "critical slab, half-thickness 9.4959 cm",
           pop=pop, active=active, inactive=inactive, seed=seed,
           flux=("  flux { type collisionClerk; map { type energyMap; grid log; min 0.001; max 20; N 300; }\n"
                 "         response (flux); flux { type fluxResponse; } }\n") if inf else "",
           bc="1 1 1 1 1 1" if inf else "0 0 1 1 1 1", hw="10.0" if inf else "9.4959", pn="P0" if inf else "P1")



# ---------------------------------------------------------------------------------------
# Continuous-energy decks (authored from the nuclides bundled with the reference; the cards travel as the binary fixtures
# data/ace/*.acebin, see tests/golden/make_ace_fixtures.py)
CE_BASE = [("1001", "1001JEF311"), ("92233", "92233JEF311"), ("52126", "52126JEF311"), ("91231", "91231JEF311"), ("91232", "91232JEF311")]


def synth_ace_cards(outdir):
    """BASELINE configs[4] (SURVEY.md section 8d cfg 5): ~20 nuclides per fuel material made from the five bundled cards by
    seeded shifts of the ESZ energy grid (interior points scaled UP by up to 2 %, end points kept, order kept; every other
    block untouched, so reaction thresholds stay covered by their law tables). Written to decks/ce/synth/ (not tracked)."""
    import struct
    import numpy as np
    here = os.path.dirname(os.path.abspath(__file__))
    src = os.path.join(here, "..", "data", "ace")
    dst = os.path.join(outdir, "ce", "synth")
    os.makedirs(dst, exist_ok=True)
    rng = np.random.default_rng(20261017)
    lib = ["! synthetic ACE library: the five bundled cards + 15 energy-shifted clones (decks/gen_decks.py)"]
    for za, fn in CE_BASE:
        raw = open(os.path.join(src, fn + ".acebin"), "rb").read()
        lib.append("%s.03c; 1; ../../../data/ace/%s.acebin;" % (za, fn))
        head = raw[:8 + 16 + 16]
        nxs = np.frombuffer(raw, np.int32, 16, 40).copy(); jxs = np.frombuffer(raw, np.int32, 32, 104).copy()
        n = struct.unpack_from("<q", raw, 232)[0]
        xss = np.frombuffer(raw, np.float64, n, 240).copy()
        for k in (10, 11, 12):
            x = xss.copy()
            nes, p0 = int(nxs[2]), int(jxs[0]) - 1
            g = x[p0:p0 + nes].copy()
            s = 1.0 + rng.uniform(0.002, 0.02)
            gi = g[1:-1] * s
            ok = gi < g[-2]                                   # keep the top of the grid as it is
            g2 = g.copy(); g2[1:-1] = np.where(ok, gi, g[1:-1])
            g2 = np.maximum.accumulate(g2)
            bad = np.where(np.diff(g2) <= 0)[0]
            for i in bad:                                      # a shifted point must not land on its neighbour
                if g[i + 1] > g[i]:
                    g2[:] = g
                    break
            x[p0:p0 + nes] = g2
            with open(os.path.join(dst, "%s_%d.acebin" % (za, k)), "wb") as f:
                zaid = ("%s.%dc" % (za, k)).encode().ljust(16, b"\0")
                f.write(raw[:8] + zaid + raw[24:40] + nxs.tobytes() + jxs.tobytes() + struct.pack("<q", n) + x.tobytes())
            lib.append("%s.%dc; 1; %s_%d.acebin;" % (za, k, za, k))
    with open(os.path.join(dst, "aceLib"), "w") as f:
        f.write("\n".join(lib) + "\n")


def ce_assembly17(pop=1000000, inactive=20, active=20, seed=20261017):
    """Synthetic CE 17x17 assembly (BASELINE configs[4]): 264 fuel pins with 20 nuclides each + 25 water holes, reflective
    boundaries, delta tracking, k-eff only."""
    guide = {(2, 5), (2, 8), (2, 11), (3, 3), (3, 13), (5, 2), (5, 5), (5, 8), (5, 11), (5, 14), (8, 2), (8, 5), (8, 8), (8, 11), (8, 14),
             (11, 2), (11, 5), (11, 8), (11, 11), (11, 14), (13, 3), (13, 13), (14, 5), (14, 8), (14, 11)}
    rows = [[2 if (r, c) in guide else 1 for c in range(17)] for r in range(17)]
    tags = ["03", "10", "11", "12"]
    fuel = []
    for za, tot in (("92233", 1.5e-4), ("52126", 2.2e-2), ("91231", 5.0e-5), ("91232", 2.0e-6), ("1001", 4.0e-6)):
        for i, t in enumerate(tags):
            fuel.append("%s.%s %.6E;" % (za, t, tot * (0.4, 0.3, 0.2, 0.1)[i]))
    water = []
    for za, tot in (("1001", 6.67e-2), ("52126", 1.0e-3)):
        for i, t in enumerate(tags):
            water.append("%s.%s %.6E;" % (za, t, tot * (0.4, 0.3, 0.2, 0.1)[i]))
    return """// Synthetic continuous-energy 17x17 assembly (BASELINE configs[4] / SURVEY.md section 8d cfg 5): 20 nuclides per fuel
// material = the five bundled nuclides + energy-shifted clones (decks/ce/synth, written by decks/gen_decks.py); unionised-grid stress.
type eigenPhysicsPackage;
pop      %(pop)d;
active   %(active)d;
inactive %(inactive)d;
seed     %(seed)d;
XSdata   ce;
dataType ce;

collisionOperator { neutronCE { type neutronCEstd; } }
transportOperator { type transportOperatorDT; }

inactiveTally { }
activeTally { }

geometry {
  type geometryStd;
  boundary (1 1 1 1 1 1);
  graph { type shrunk; }
  surfaces { bound { id 1; type box; origin (0.0 0.0 0.0); halfwidth (10.71 10.71 10.0); } }
  cells { }
  universes {
    root { id 100; type rootUniverse; border 1; fill u<10>; }
    pin  { id 1; type pinUniverse; radii (0.41 0.0); fills (fuel water); }
    hole { id 2; type pinUniverse; radii (0.0); fills (water); }
    lat  { id 10; type latUniverse; origin (0.0 0.0 0.0); pitch (1.26 1.26 0.0); shape (17 17 0); padMat water;
      map (
%(map)s ); }
  }
}

nuclearData {
  handles { ce { type aceNeutronDatabase; aceLibrary ./synth/aceLib; ures 0; majorant 1; } }
  materials {
    fuel  { temp 293; composition { %(fuel)s } }
    water { temp 293; composition { %(water)s } }
  }
}
""" % dict(pop=pop, inactive=inactive, active=active, seed=seed, map=fmt_map(rows), fuel=" ".join(fuel), water=" ".join(water))


def fixed_mg(pop=100000, cycles=50, seed=20261017):
    """Fixed-source multigroup problem (fixedSourcePhysicsPackage): isotropic group-1 point source in a UO2 sphere inside a water
    box with vacuum boundaries (subcritical: fission neutrons are followed as secondaries of the same history)."""
    return """// fixed-source multigroup: point source in a small UO2 sphere (C5G7 constants), vacuum box
type fixedSourcePhysicsPackage;
pop      %(pop)d;
cycles   %(cycles)d;
seed     %(seed)d;
XSdata   mg;
dataType mg;
buffer   50;

collisionOperator { neutronMG { type neutronMGstd; } }
transportOperator { type transportOperatorDT; }

source { type pointSource; r (0.1 0.2 0.3); particle neutron; G 1; }

tally {
  norm fiss; normVal 100;
  fiss { type collisionClerk; response (fiss); fiss { type macroResponse; MT -6; } }
  flux { type collisionClerk; map { type spaceMap; axis x; grid lin; min -5.0; max 5.0; N 20; } response (flux abs); flux { type fluxResponse; } abs { type macroResponse; MT -2; } }
}

geometry {
  type geometryStd;
  boundary (0 0 0 0 0 0);
  graph { type shrunk; }
  surfaces {
    ball  { id 1; type sphere; origin (0.0 0.0 0.0); radius 3.0; }
    bound { id 2; type box; origin (0.0 0.0 0.0); halfwidth (5.0 5.0 5.0); }
  }
  cells {
    in  { id 1; type simpleCell; surfaces (-1); filltype mat; material UO2; }
    out { id 2; type simpleCell; surfaces (1);  filltype mat; material water; }
  }
  universes {
    root { id 1; type rootUniverse; border 2; fill u<2>; }
    geom { id 2; type cellUniverse; cells (1 2); }
  }
}

nuclearData {
  handles { mg { type baseMgNeutronDatabase; PN P0; } }
  materials {
    UO2   { temp 300; xsFile ../c5g7/xs/UO2.xs; composition { } }
    water { temp 300; xsFile ../c5g7/xs/moder.xs; composition { } }
  }
}
""" % dict(pop=pop, cycles=cycles, seed=seed)


def fixed_ce(pop=100000, cycles=20, seed=20261017):
    """Fixed-source continuous-energy problem in the spirit of the reference's InputFiles/sphere_with_DT: 14.1 MeV point source in
    a hollow fuel sphere (the bundled nuclides), vacuum boundary, surface tracking."""
    return """// fixed-source continuous energy: 14.1 MeV point source in a fuel sphere (InputFiles/sphere_with_DT with the bundled nuclides)
type fixedSourcePhysicsPackage;
pop      %(pop)d;
cycles   %(cycles)d;
seed     %(seed)d;
XSdata   ce;
dataType ce;

collisionOperator { neutronCE { type neutronCEstd; } }
transportOperator { type transportOperatorST; }

source { type pointSource; r (0.0 0.0 0.0); particle neutron; E 14.1; }

tally {
  fiss { type collisionClerk; response (fiss); fiss { type macroResponse; MT -6; } }
  flux { type collisionClerk;
         map { type multiMap; maps (ene mat);
               ene { type energyMap; grid log; min 0.001; max 20.0; N 100; }
               mat { type materialMap; materials (fuel water); } }
         response (flux); flux { type fluxResponse; } }
}

geometry {
  type geometryStd;
  boundary (0 0 0 0 0 0);
  graph { type shrunk; }
  surfaces {
    hollow { id 1; type sphere; origin (0.0 0.0 0.0); radius 0.5; }
    fuel   { id 2; type sphere; origin (0.0 0.0 0.0); radius 4.0; }
    bound  { id 3; type sphere; origin (0.0 0.0 0.0); radius 6.0; }
  }
  cells {
    inside { id 1; type simpleCell; surfaces (-1);   filltype mat; material water; }
    fuel   { id 2; type simpleCell; surfaces (1 -2); filltype mat; material fuel; }
    refl   { id 3; type simpleCell; surfaces (2);    filltype mat; material water; }
  }
  universes {
    root { id 1; type rootUniverse; border 3; fill u<2>; }
    geom { id 2; type cellUniverse; cells (1 2 3); }
  }
}

nuclearData {
  handles { ce { type aceNeutronDatabase; aceLibrary ../../data/ace/aceLib; ures 0; majorant 1; } }
  materials {
    fuel  { temp 293; composition { 92233.03 1.0E-3; 52126.03 2.2E-2; 91231.03 5.0E-5; 91232.03 2.0E-6; } }
    water { temp 293; composition { 1001.03 6.67E-2; 52126.03 1.0E-3; } }
  }
}
""" % dict(pop=pop, cycles=cycles, seed=seed)


def write(outdir):
    def put(rel, text):
        p = os.path.join(outdir, rel)
        os.makedirs(os.path.dirname(p), exist_ok=True)
        with open(p, "w") as f:
            f.write(text)

    for name in C5G7_XS:
        put("c5g7/xs/%s.xs" % name, xs_text(name))
    put("c5g7/c5g7_2d", c5g7_2d())
    put("c5g7/c5g7_2d_ht", c5g7_2d(tracking="HT"))
    put("c5g7/c5g7_3d_rodded", c5g7_3d_rodded())
    put("urr/xs/URRa_2_1.xs", "// Sood et al. URRa-2-1 two-group constants\n" + URR_2G)
    put("urr/inf", urr_deck("inf"))
    put("urr/slab", urr_deck("slab"))
    synth_ace_cards(outdir)
    put("ce/assembly17", ce_assembly17())
    put("fixed/mg_sphere", fixed_mg())
    put("fixed/ce_sphere", fixed_ce())


if __name__ == "__main__":
    write(sys.argv[1] if len(sys.argv) > 1 else os.path.dirname(os.path.abspath(__file__)))
