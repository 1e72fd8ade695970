"""The input-dictionary parser shared by the host package and the oracle (scone_b200/csrc/host/dict.hpp) against the reference's own
parser tests: DataStructures/Tests/dictParser_test.f90 (charToDict) and dictParser_iTest.f90 (IntegrationTestFiles/testDictionary)."""
import ctypes as C

import pytest

import scone_b200.lib as sl

TAPE = (" myInt 7;  myChar my;  myReal 1.3;  weirdFloat 1E-11;  intArray (1 2 4 5);  realArray (1.1 2.2 3.4 1E-11); "
        " charArray (One element );  subDict { myInt 3; myReal 3.2; }")
# IntegrationTestFiles/testDictionary (input data): odd spacing, the three comment markers, a sub-dictionary glued to its name
FILE_TEXT = """myInt 7;
myChar my;
myReal 1.3         ;
intArray (1 2 4 5);
realArray (      1
               2.2  3.5 )   ; // a bit of weird formatting
charArray (One element );
!Another comment
subDict{ myInt 3; myReal 3.2; } ! Note lack of space

//Line comment without space after marker. This case caused a bug at some point
subDict2 { myInt 4; myReal 17.0;}
"""


def get(text, path, kind, is_path=False):
    L = sl.load_library()
    out = C.create_string_buffer(1 << 16)
    n = L.sbh_dict_get(text.encode(), 1 if is_path else 0, path.encode(), kind.encode(), out, 1 << 16)
    if n < 0:
        raise ValueError(L.sbh_last_error(None).decode())
    return out.value.decode()


def test_char_to_dict_known_answers():
    # dictParser_test.f90:17-32
    assert get(TAPE, "myInt", "i") == "7" and get(TAPE, "intArray", "I") == "1 2 4 5"
    assert float(get(TAPE, "myReal", "r")) == 1.3
    assert [float(x) for x in get(TAPE, "realArray", "R").split()] == [1.1, 2.2, 3.4, 1.0e-11]
    assert float(get(TAPE, "weirdFloat", "r")) == 1.0e-11                 # a real without a decimal point
    assert get(TAPE, "myChar", "w") == "my" and get(TAPE, "charArray", "W") == "One element"
    assert get(TAPE, "subDict/myInt", "i") == "3" and float(get(TAPE, "subDict/myReal", "r")) == 3.2
    assert get(TAPE, "", "k").split() == ["myInt", "myChar", "myReal", "weirdFloat", "intArray", "realArray", "charArray", "subDict"]


def test_file_to_dict_known_answers(tmp_path):
    # dictParser_iTest.f90:18-37
    p = tmp_path / "testDictionary"
    p.write_text(FILE_TEXT)
    path = str(p)
    assert get(path, "myInt", "i", True) == "7" and get(path, "intArray", "I", True) == "1 2 4 5"
    assert float(get(path, "myReal", "r", True)) == 1.3
    assert [float(x) for x in get(path, "realArray", "R", True).split()] == [1.0, 2.2, 3.5]      # an integer inside a real array
    assert get(path, "subDict/myInt", "i", True) == "3" and float(get(path, "subDict/myReal", "r", True)) == 3.2
    assert get(path, "subDict2/myInt", "i", True) == "4" and float(get(path, "subDict2/myReal", "r", True)) == 17.0


def test_number_tokens_follow_the_fortran_reads():
    """convert_reader (dictParser_func.f90:766-793): I20 read first, then ES100.0, else a word."""
    t = "a 1e-3; b 2.5+3; c 1.0d2; d 12345678901; e .5; f 5.; g -7; h +3; i 1.5x; j e5; k 3-2; l (1 2.0 3);"
    for key, ref in (("a", 1e-3), ("b", 2500.0), ("c", 100.0), ("d", 12345678901.0), ("e", 0.5), ("f", 5.0), ("g", -7.0), ("h", 3.0), ("k", 0.03)):
        assert float(get(t, key, "r")) == ref
    assert get(t, "g", "i") == "-7" and get(t, "h", "i") == "3"
    for key in ("i", "j"):
        with pytest.raises(ValueError):
            get(t, key, "r")
        assert get(t, key, "w") in ("1.5x", "e5")
    with pytest.raises(ValueError):
        get(t, "d", "i")                                                  # does not fit shortInt: it is a real
    with pytest.raises(ValueError):
        get(t, "l", "I")                                                  # a list with a real in it is a real list
    assert [float(x) for x in get(t, "l", "R").split()] == [1.0, 2.0, 3.0]
