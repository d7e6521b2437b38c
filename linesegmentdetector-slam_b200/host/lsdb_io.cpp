// The reference's text formats, read and written on the host (SURVEY.md §8b "file formats to keep"); part of liblsdb200.so.
//   mapParam.txt  : `cols rows resol oriX oriY`                                   (LSD/main_on_windows.cpp:28-34)
//   mapValue*.txt : rows x cols integers, stored through %d into a uint8 slot,
//                   i.e. value & 0xFF (-1 -> 255)                                 (LSD/main_on_windows.cpp:38-46)
//   mapCache.txt  : rows x cols doubles, whitespace separated, row-major          (LSD/test.cpp:11-17)
// No device is involved; errors are LSDB_ERR_ARG (bad argument / unreadable file / short file).
#include "../../include/lsdb200.h"

#include <stdio.h>
#include <stdlib.h>

extern "C" int lsdb_read_map_param(const char* path, int* cols, int* rows, double* resol, double* ori_x, double* ori_y) {
    if (!path || !cols || !rows || !resol || !ori_x || !ori_y) return LSDB_ERR_ARG;
    FILE* fp = fopen(path, "r");
    if (!fp) return LSDB_ERR_ARG;
    const int n = fscanf(fp, "%d %d %lf %lf %lf", cols, rows, resol, ori_x, ori_y);
    fclose(fp);
    return n == 5 ? LSDB_OK : LSDB_ERR_ARG;
}

extern "C" int lsdb_read_map_value(const char* path, int cols, int rows, uint8_t* out) {
    if (!path || !out || cols <= 0 || rows <= 0) return LSDB_ERR_ARG;
    FILE* fp = fopen(path, "r");
    if (!fp) return LSDB_ERR_ARG;
    const size_t n = (size_t)cols * rows;
    for (size_t i = 0; i < n; i++) {
        int v;
        if (fscanf(fp, "%d", &v) != 1) { fclose(fp); return LSDB_ERR_ARG; }
        out[i] = (uint8_t)(v & 0xFF);   // what the reference's `%d` into a uint8_t* leaves in the pixel
    }
    fclose(fp);
    return LSDB_OK;
}

extern "C" int lsdb_read_map_cache(const char* path, int cols, int rows, double* out) {
    if (!path || !out || cols <= 0 || rows <= 0) return LSDB_ERR_ARG;
    FILE* fp = fopen(path, "r");
    if (!fp) return LSDB_ERR_ARG;
    const size_t n = (size_t)cols * rows;
    for (size_t i = 0; i < n; i++)
        if (fscanf(fp, "%lf", &out[i]) != 1) { fclose(fp); return LSDB_ERR_ARG; }
    fclose(fp);
    return LSDB_OK;
}

// %.17g round-trips every double through the reference's `%lf` reader
extern "C" int lsdb_write_map_cache(const char* path, int cols, int rows, const double* in) {
    if (!path || !in || cols <= 0 || rows <= 0) return LSDB_ERR_ARG;
    FILE* fp = fopen(path, "w");
    if (!fp) return LSDB_ERR_ARG;
    for (int y = 0; y < rows; y++) {
        for (int x = 0; x < cols; x++)
            if (fprintf(fp, x + 1 < cols ? "%.17g " : "%.17g\n", in[(size_t)y * cols + x]) < 0) { fclose(fp); return LSDB_ERR_ARG; }
    }
    return fclose(fp) == 0 ? LSDB_OK : LSDB_ERR_ARG;
}
