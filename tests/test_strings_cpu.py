"""Oracle pin for string keys: the byte order the dictionary codes must preserve, against the reference's own vectors
(common/unsafe/src/test/java/org/apache/spark/unsafe/types/UTF8StringSuite.java:100-112, binaryCompareTo)."""
import pyarrow as pa

from oracle import oracle as O

# (a, b, sign of a.binaryCompare(b)) -- UTF8StringSuite.java:101-111
BINARY_COMPARE_VECTORS = [
    ("", "a", -1), ("abc", "ABC", 1), ("abc0", "abc", 1), ("abcabcabc", "abcabcabc", 0), ("aBcabcabc", "Abcabcabc", 1),
    ("Abcabcabc", "abcabcabC", -1), ("abcabcabc", "abcabcabC", 1), ("abc", "世界", -1), ("你好", "世界", 1), ("你好123", "你好122", 1),
]


def sign(x):
    return (x > 0) - (x < 0)


def test_binary_compare_matches_the_reference_vectors():
    for a, b, want in BINARY_COMPARE_VECTORS:
        assert sign(O.binary_compare(a.encode(), b.encode())) == want, (a, b)
        assert sign(O.binary_compare(b.encode(), a.encode())) == -want


def test_codes_preserve_equality_and_order():
    words = sorted({w for a, b, _ in BINARY_COMPARE_VECTORS for w in (a, b)})
    col = pa.chunked_array([pa.array(words + [None] + words[::-1])])
    codes, dictionary = O.string_codes(col)
    c = codes.to_pylist()
    v = col.to_pylist()
    for i in range(len(v)):
        for j in range(len(v)):
            if v[i] is None or v[j] is None:
                assert (c[i] is None) == (v[i] is None)
                continue
            assert sign(c[i] - c[j]) == sign(O.binary_compare(v[i].encode(), v[j].encode()))
    assert [d.decode() for d in dictionary] == [dictionary[k].decode() for k in range(len(dictionary))]
    t = pa.table({"s": col, "x": list(range(len(v)))})
    enc, dicts = O.encode_string_columns(t, ["s"])
    back = O.decode_string_columns(enc, dicts)
    assert back.column("s").to_pylist() == v
