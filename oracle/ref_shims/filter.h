// TEST INFRASTRUCTURE (oracle/_ref build only). Stand-in for the reference's include/filter.h, whose real version drags
// rocksdb/store headers that are absent here. src/or_iterator.cpp:2 only includes it; include/art.h:281-285 and
// src/art.cpp name the comparator enumeration (include/filter.h:16-26) in the numeric-search signatures, so its
// enumerators are declared here in the reference's order. Nothing else from it is used by the path.
#pragma once
enum NUM_COMPARATOR { LESS_THAN, LESS_THAN_EQUALS, EQUALS, NOT_EQUALS, CONTAINS, GREATER_THAN, GREATER_THAN_EQUALS, RANGE_INCLUSIVE,
                      CONTAINS_PHRASE };
