// TEST INFRASTRUCTURE (oracle/_ref build only). Empty stand-in: the reference's src/or_iterator.cpp:2 includes
// "filter.h", whose real version drags rocksdb/store headers that are absent here. Nothing from it is used
// by the posting-list path.
#pragma once
