// Stub of Assimp for the oracle build ONLY (inc/default_schema.hpp:13-15,516-545 must parse).
// ReadFile always returns nullptr; the oracle never loads meshes through the reference loader.
#ifndef ORACLE_STUB_ASSIMP_IMPORTER_HPP
#define ORACLE_STUB_ASSIMP_IMPORTER_HPP
struct aiVector3D { float x, y, z; };
struct aiFace { unsigned int mNumIndices; unsigned int *mIndices; };
struct aiMesh { unsigned int mPrimitiveTypes; unsigned int mNumFaces; aiFace *mFaces; aiVector3D *mVertices; };
struct aiScene { unsigned int mNumMeshes; aiMesh **mMeshes; };
enum { aiPrimitiveType_TRIANGLE = 0x4 };
enum { aiProcess_CalcTangentSpace = 0x1, aiProcess_JoinIdenticalVertices = 0x2, aiProcess_Triangulate = 0x8, aiProcess_SortByPType = 0x8000 };
namespace Assimp {
class Importer {
public:
  const aiScene *ReadFile(const char *, unsigned int) { return nullptr; }
};
}
#endif
