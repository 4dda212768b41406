// Declaration-only stand-in so that the reference's ColladaLoader.h (included by Mesh.h) parses.
// pugixml is not vendored by the reference and ColladaLoader.cpp is NOT compiled into oracle/_ref.
#pragma once
namespace pugi {
class xml_node {};
class xml_document {};
}
