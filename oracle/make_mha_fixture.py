"""Writes tests/golden/itk_style_*.mha: MetaImage files laid out byte for byte the way ITK's MetaImageIO (what
SimpleITK's ImageFileWriter uses, reference utils.py:87-104 / dataset.py:49-55) lays them out.  TEST INFRASTRUCTURE.

SimpleITK is not installed in this image, so the files are produced here from the MetaIO conventions rather than by
ITK itself; this script deliberately shares no code with dram_b200/mha_io.py (the reader under test):
  * header keys in MetaIO's fixed order: ObjectType, NDims, BinaryData, BinaryDataByteOrderMSB, CompressedData,
    [CompressedDataSize], TransformMatrix, Offset, CenterOfRotation, AnatomicalOrientation, ElementSpacing, DimSize,
    ElementType, ElementDataFile = LOCAL, then the voxels (x fastest), zlib-deflated when CompressedData = True;
  * doubles printed with 17 significant digits (MetaIO's SetDoublePrecision(17): 0.7 -> 0.69999999999999996);
  * TransformMatrix row i = direction cosines of image axis i, i.e. the TRANSPOSE of the row-major matrix
    itk::Image::GetDirection() / sitk GetDirection() returns (itkMetaImageIO.cxx writes GetDirection(i)[j] at [i*n+j]);
  * AnatomicalOrientation = RAI for an identity-like direction, ??? when ITK cannot name it.
The JSON next to the files holds what sitk.ReadImage(...) would report (GetSpacing/GetOrigin/GetDirection, x-y-z order)
and the voxel formula.

    python -m oracle.make_mha_fixture
"""
import json
import os
import struct
import zlib

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def g17(v):
    return "%.17g" % v


def voxel(x, y, z, kind):
    if kind == "int16":
        return ((x * 37 + y * 101 + z * 977) % 4096) - 1024
    return (x * 7 + y * 13 + z * 29) % 6  # lobe labels 0..5


def build(kind, dims, spacing, origin, direction_rowmajor, compressed, msb, orientation):
    nx, ny, nz = dims
    fmt = (">" if msb else "<") + ("h" if kind == "int16" else "B")
    raw = b"".join(struct.pack(fmt, voxel(x, y, z, kind)) for z in range(nz) for y in range(ny) for x in range(nx))
    payload = zlib.compress(raw, 2) if compressed else raw
    d = direction_rowmajor
    transform = [d[0], d[3], d[6], d[1], d[4], d[7], d[2], d[5], d[8]]  # transpose
    lines = ["ObjectType = Image", "NDims = 3", "BinaryData = True",
             f"BinaryDataByteOrderMSB = {'True' if msb else 'False'}", f"CompressedData = {'True' if compressed else 'False'}"]
    if compressed:
        lines.append(f"CompressedDataSize = {len(payload)}")
    lines += ["TransformMatrix = " + " ".join(g17(v) for v in transform), "Offset = " + " ".join(g17(v) for v in origin),
              "CenterOfRotation = 0 0 0", f"AnatomicalOrientation = {orientation}",
              "ElementSpacing = " + " ".join(g17(v) for v in spacing), f"DimSize = {nx} {ny} {nz}",
              f"ElementType = {'MET_SHORT' if kind == 'int16' else 'MET_UCHAR'}", "ElementDataFile = LOCAL"]
    return ("\n".join(lines) + "\n").encode("ascii") + payload


def main():
    import math

    c, s = math.cos(math.radians(12.0)), math.sin(math.radians(12.0))
    cases = {
        # an oblique acquisition (gantry tilt about x), compressed, as ImageFileWriter.SetUseCompression(True) writes
        "itk_style_ct_oblique": dict(kind="int16", dims=(7, 6, 5), spacing=(0.7, 0.7, 1.25), origin=(-180.5, -150.0, -300.25),
                                     direction_rowmajor=(1.0, 0.0, 0.0, 0.0, c, -s, 0.0, s, c), compressed=True, msb=False,
                                     orientation="???"),
        # a lobe mask, identity direction, uncompressed, big-endian flag set (no effect on 1-byte voxels)
        "itk_style_lobes_plain": dict(kind="uint8", dims=(7, 6, 5), spacing=(0.7, 0.7, 1.25), origin=(-180.5, -150.0, -300.25),
                                      direction_rowmajor=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0), compressed=False,
                                      msb=True, orientation="RAI"),
        # big-endian 16-bit voxels (old scanners' exports)
        "itk_style_ct_msb": dict(kind="int16", dims=(4, 3, 2), spacing=(1.0, 1.0, 2.5), origin=(0.0, 0.0, 0.0),
                                 direction_rowmajor=(1.0, 0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 1.0), compressed=False, msb=True,
                                 orientation="RAI"),
    }
    expect = {}
    for name, c_ in cases.items():
        with open(os.path.join(GOLDEN, name + ".mha"), "wb") as f:
            f.write(build(**c_))
        expect[name] = {"kind": c_["kind"], "dims_xyz": list(c_["dims"]), "spacing": list(c_["spacing"]),
                        "origin": list(c_["origin"]), "direction": list(c_["direction_rowmajor"])}
    with open(os.path.join(GOLDEN, "itk_style_mha.json"), "w") as f:
        json.dump(expect, f, indent=1)
    print("wrote", sorted(expect))


if __name__ == "__main__":
    main()
