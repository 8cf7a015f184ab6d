"""python-protobuf message classes for the reference's meta.pb — an INDEPENDENT parser/serialiser for the pins.

The descriptor below mirrors point_viewer_proto_rust/src/proto.proto:44-149 (package point_viewer.proto, proto3):
field names, numbers, types and the `oneof data` of `Meta` are restated by hand; google.protobuf does all wire-format
work.  Nothing of this repository's own encoder/decoder (csrc/disk_io.hpp, oracle/oracle_disk.hpp) is involved, which is
the point: both are checked against it (tests/test_meta_protobuf.py).
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, unknown_fields

F = descriptor_pb2.FieldDescriptorProto


def _field(msg, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None, oneof=None):
    f = msg.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = ".point_viewer.proto." + type_name
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _build_pool():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "point_viewer_meta_restated.proto"
    fd.package = "point_viewer.proto"
    fd.syntax = "proto3"

    m = fd.message_type.add()
    m.name = "Vector3f"  # proto.proto:27-31
    for i, n in enumerate("xyz"):
        _field(m, n, i + 1, F.TYPE_FLOAT)
    m = fd.message_type.add()
    m.name = "Vector3d"  # proto.proto:33-37
    for i, n in enumerate("xyz"):
        _field(m, n, i + 1, F.TYPE_DOUBLE)

    m = fd.message_type.add()
    m.name = "AxisAlignedCuboid"  # proto.proto:58-66
    _field(m, "min", 3, F.TYPE_MESSAGE, type_name="Vector3d")
    _field(m, "max", 4, F.TYPE_MESSAGE, type_name="Vector3d")
    _field(m, "deprecated_min", 1, F.TYPE_MESSAGE, type_name="Vector3f")
    _field(m, "deprecated_max", 2, F.TYPE_MESSAGE, type_name="Vector3f")

    m = fd.message_type.add()
    m.name = "NodeId"  # proto.proto:68-76
    _field(m, "high", 3, F.TYPE_UINT64)
    _field(m, "low", 4, F.TYPE_UINT64)
    _field(m, "deprecated_level", 1, F.TYPE_INT32)
    _field(m, "deprecated_index", 2, F.TYPE_INT64)

    e = fd.enum_type.add()
    e.name = "PositionEncoding"  # proto.proto:78-84
    for n, v in (("INVALID", 0), ("Uint8", 1), ("Uint16", 2), ("Float32", 3), ("Float64", 4)):
        ev = e.value.add()
        ev.name, ev.number = n, v

    m = fd.message_type.add()
    m.name = "OctreeNode"  # proto.proto:86-90
    _field(m, "position_encoding", 2, F.TYPE_ENUM, type_name="PositionEncoding")
    _field(m, "num_points", 3, F.TYPE_INT64)
    _field(m, "id", 4, F.TYPE_MESSAGE, type_name="NodeId")

    e = fd.enum_type.add()
    e.name = "AttributeDataType"  # proto.proto:92-111
    for n, v in (("INVALID_DATA_TYPE", 0), ("U8", 1), ("U16", 2), ("U32", 3), ("U64", 4), ("I8", 6), ("I16", 7), ("I32", 8), ("I64", 9), ("F32", 11),
                 ("F64", 12), ("U8Vec3", 27), ("F64Vec3", 38)):
        ev = e.value.add()
        ev.name, ev.number = n, v
    m = fd.message_type.add()
    m.name = "Attribute"  # proto.proto:113-116
    _field(m, "name", 1, F.TYPE_STRING)
    _field(m, "data_type", 2, F.TYPE_ENUM, type_name="AttributeDataType")
    m = fd.message_type.add()
    m.name = "S2Cell"  # proto.proto:118-121
    _field(m, "id", 1, F.TYPE_UINT64)
    _field(m, "num_points", 2, F.TYPE_UINT64)

    m = fd.message_type.add()
    m.name = "OctreeMeta"  # proto.proto:123-129
    _field(m, "resolution", 2, F.TYPE_DOUBLE)
    _field(m, "nodes", 3, F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name="OctreeNode")
    _field(m, "deprecated_bounding_box", 1, F.TYPE_MESSAGE, type_name="AxisAlignedCuboid")
    m = fd.message_type.add()
    m.name = "S2Meta"  # proto.proto:131-134
    _field(m, "cells", 1, F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name="S2Cell")
    _field(m, "attributes", 2, F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name="Attribute")

    m = fd.message_type.add()
    m.name = "Meta"  # proto.proto:137-149
    m.oneof_decl.add().name = "data"
    _field(m, "version", 1, F.TYPE_INT32)
    _field(m, "bounding_box", 4, F.TYPE_MESSAGE, type_name="AxisAlignedCuboid")
    _field(m, "octree", 6, F.TYPE_MESSAGE, type_name="OctreeMeta", oneof=0)
    _field(m, "s2", 7, F.TYPE_MESSAGE, type_name="S2Meta", oneof=0)
    _field(m, "deprecated_resolution", 3, F.TYPE_DOUBLE)
    _field(m, "deprecated_nodes", 5, F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name="OctreeNode")

    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


_POOL = _build_pool()
Meta = message_factory.GetMessageClass(_POOL.FindMessageTypeByName("point_viewer.proto.Meta"))


def _count_unknown(m):
    """Unknown fields anywhere in the message tree (a writer inventing field numbers would show up here)."""
    total = len(unknown_fields.UnknownFieldSet(m))
    for fd, v in m.ListFields():
        if fd.type == fd.TYPE_MESSAGE:
            for sub in (v if fd.is_repeated else [v]):
                total += _count_unknown(sub)
    return total


def parse_meta(data):
    """bytes of a meta.pb -> dict(version, bbox_min, bbox_max, resolution, nodes={(high, low): (num_points, enc)})."""
    m = Meta()
    m.ParseFromString(data)
    assert m.WhichOneof("data") == "octree"
    bb = m.bounding_box
    nodes = {}
    for n in m.octree.nodes:
        key = (int(n.id.high), int(n.id.low))
        assert key not in nodes, "duplicate node id"
        nodes[key] = (int(n.num_points), int(n.position_encoding))
    return dict(version=int(m.version), bbox_min=(bb.min.x, bb.min.y, bb.min.z), bbox_max=(bb.max.x, bb.max.y, bb.max.z),
                resolution=m.octree.resolution, nodes=nodes, unknown=_count_unknown(m))


def serialize_meta(bbox_min, bbox_max, resolution, nodes, version=13):
    """nodes: iterable of (high, low, num_points, enc) -> meta.pb bytes as python-protobuf writes them (to_meta_proto, octree/mod.rs:87-99)."""
    m = Meta()
    m.version = version
    for a, v in zip("xyz", bbox_min):
        setattr(m.bounding_box.min, a, float(v))
    for a, v in zip("xyz", bbox_max):
        setattr(m.bounding_box.max, a, float(v))
    m.octree.resolution = float(resolution)
    for hi, lo, n, enc in nodes:
        p = m.octree.nodes.add()
        p.id.high, p.id.low, p.num_points, p.position_encoding = int(hi), int(lo), int(n), int(enc)
    return m.SerializeToString()


def _build_xray_pool():
    """xray_proto_rust/src/proto.proto:20-56 (package xray.proto), restated as a descriptor."""
    from google.protobuf import descriptor_pb2

    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "xray_restated.proto"
    fd.package = "xray.proto"
    fd.syntax = "proto3"

    def field(m, name, number, ftype, label=F.LABEL_OPTIONAL, type_name=None):
        f = m.field.add()
        f.name, f.number, f.type, f.label = name, number, ftype, label
        if type_name:
            f.type_name = ".xray.proto." + type_name

    m = fd.message_type.add()
    m.name = "Vector2f"
    field(m, "x", 1, F.TYPE_FLOAT)
    field(m, "y", 2, F.TYPE_FLOAT)
    m = fd.message_type.add()
    m.name = "Vector2d"
    field(m, "x", 1, F.TYPE_DOUBLE)
    field(m, "y", 2, F.TYPE_DOUBLE)
    m = fd.message_type.add()
    m.name = "Rect"
    field(m, "min", 3, F.TYPE_MESSAGE, type_name="Vector2d")
    field(m, "edge_length", 4, F.TYPE_DOUBLE)
    field(m, "deprecated_min", 1, F.TYPE_MESSAGE, type_name="Vector2f")
    field(m, "deprecated_edge_length", 2, F.TYPE_FLOAT)
    m = fd.message_type.add()
    m.name = "NodeId"
    field(m, "level", 1, F.TYPE_UINT32)
    field(m, "index", 2, F.TYPE_UINT64)
    m = fd.message_type.add()
    m.name = "Meta"
    field(m, "version", 1, F.TYPE_INT32)
    field(m, "bounding_rect", 2, F.TYPE_MESSAGE, type_name="Rect")
    field(m, "deepest_level", 3, F.TYPE_UINT32)
    field(m, "tile_size", 4, F.TYPE_UINT32)
    field(m, "nodes", 5, F.TYPE_MESSAGE, label=F.LABEL_REPEATED, type_name="NodeId")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


XrayMeta = message_factory.GetMessageClass(_build_xray_pool().FindMessageTypeByName("xray.proto.Meta"))
