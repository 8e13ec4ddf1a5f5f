// Stub: declaration only (inc/images.hpp includes it); never called by the oracle.
#ifndef ORACLE_STUB_STB_IMAGE_WRITE_H
#define ORACLE_STUB_STB_IMAGE_WRITE_H
extern "C" int stbi_write_jpg(char const *filename, int x, int y, int comp, const void *data, int quality);
#endif
