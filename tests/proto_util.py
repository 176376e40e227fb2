"""nerrf.trace message classes built at run time with the protobuf RUNTIME (no protoc in this image): the
independent, canonical implementation the ingest decoder is pinned against.  Field numbers / types follow
proto/trace.proto:11-49 of the reference."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory, timestamp_pb2  # noqa: F401

_CLASSES = None


def classes():
    global _CLASSES
    if _CLASSES is not None:
        return _CLASSES
    F = descriptor_pb2.FieldDescriptorProto
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name = "nerrf_trace_for_tests.proto"; fd.package = "nerrf.trace"; fd.syntax = "proto3"
    fd.dependency.append("google/protobuf/timestamp.proto")
    m = fd.message_type.add(); m.name = "Event"

    def add(name, num, typ, label=F.LABEL_OPTIONAL, type_name=None):
        f = m.field.add(); f.name = name; f.number = num; f.type = typ; f.label = label
        if type_name:
            f.type_name = type_name

    add("ts", 1, F.TYPE_MESSAGE, type_name=".google.protobuf.Timestamp")
    add("pid", 2, F.TYPE_UINT32); add("tid", 3, F.TYPE_UINT32); add("comm", 4, F.TYPE_STRING)
    add("syscall", 5, F.TYPE_STRING); add("path", 6, F.TYPE_STRING); add("new_path", 7, F.TYPE_STRING)
    e = m.enum_type.add(); e.name = "OpenFlags"
    for n, v in (("O_RDONLY", 0), ("O_WRONLY", 1), ("O_RDWR", 2)):
        x = e.value.add(); x.name = n; x.number = v
    add("flags", 8, F.TYPE_ENUM, type_name=".nerrf.trace.Event.OpenFlags")
    add("ret_val", 9, F.TYPE_SINT64); add("bytes", 10, F.TYPE_UINT64); add("inode", 11, F.TYPE_STRING)
    add("mode", 12, F.TYPE_UINT32); add("uid", 13, F.TYPE_UINT64); add("gid", 14, F.TYPE_UINT64)
    add("dependencies", 15, F.TYPE_STRING, label=F.LABEL_REPEATED)
    b = fd.message_type.add(); b.name = "EventBatch"
    f = b.field.add(); f.name = "events"; f.number = 1; f.type = F.TYPE_MESSAGE; f.label = F.LABEL_REPEATED
    f.type_name = ".nerrf.trace.Event"
    pool = descriptor_pool.DescriptorPool()
    pool.AddSerializedFile(timestamp_pb2.DESCRIPTOR.serialized_pb)
    pool.AddSerializedFile(fd.SerializeToString())
    _CLASSES = (message_factory.GetMessageClass(pool.FindMessageTypeByName("nerrf.trace.Event")),
                message_factory.GetMessageClass(pool.FindMessageTypeByName("nerrf.trace.EventBatch")))
    return _CLASSES
