// VoxelLoader.cpp — MagicaVoxel .vox -> Grid (reference src/VoxelLoader.cpp:47-100).
// Asset I/O is outside the hot-path scope; this is a small self-contained reader of the public
// .vox chunk format (first model only: SIZE, XYZI, RGBA) so that init_grid_from_magika_voxel and
// the wrapper's init_grid_magikavoxel / create_grid keep working without third-party code.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include <vector>

#include "lustrine/Lustrine.hpp"

namespace Lustrine {

namespace {
struct VoxModel {
    int sx = 0, sy = 0, sz = 0;
    std::vector<uint8_t> voxels;  // index x + y*sx + z*sx*sy, 0 = empty, else palette index
    uint8_t palette[256][4];
    bool ok = false;
};

uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

VoxModel parse_vox(const std::vector<uint8_t>& buf) {
    VoxModel m;
    for (int i = 0; i < 256; i++) { m.palette[i][0] = m.palette[i][1] = m.palette[i][2] = m.palette[i][3] = 255; }
    if (buf.size() < 20 || std::memcmp(buf.data(), "VOX ", 4) != 0) return m;
    size_t pos = 8;  // magic + version
    bool have_size = false, have_voxels = false;
    // MAIN chunk header: id, content size, children size; its children follow directly
    if (std::memcmp(buf.data() + pos, "MAIN", 4) == 0) pos += 12 + rd32(buf.data() + pos + 4);
    while (pos + 12 <= buf.size()) {
        const uint8_t* h = buf.data() + pos;
        uint32_t content = rd32(h + 4), children = rd32(h + 8);
        const uint8_t* c = h + 12;
        if (pos + 12 + content > buf.size()) break;
        if (std::memcmp(h, "SIZE", 4) == 0 && !have_size && content >= 12) {
            m.sx = (int)rd32(c); m.sy = (int)rd32(c + 4); m.sz = (int)rd32(c + 8);
            m.voxels.assign((size_t)m.sx * m.sy * m.sz, 0);
            have_size = true;
        } else if (std::memcmp(h, "XYZI", 4) == 0 && have_size && !have_voxels && content >= 4) {
            uint32_t n = rd32(c);
            for (uint32_t i = 0; i < n && 4 + 4 * (size_t)(i + 1) <= content; i++) {
                const uint8_t* v = c + 4 + 4 * (size_t)i;
                if (v[0] < m.sx && v[1] < m.sy && v[2] < m.sz) m.voxels[v[0] + (size_t)v[1] * m.sx + (size_t)v[2] * m.sx * m.sy] = v[3];
            }
            have_voxels = true;
        } else if (std::memcmp(h, "RGBA", 4) == 0 && content >= 1024) {
            for (int i = 0; i < 255; i++) std::memcpy(m.palette[i + 1], c + 4 * i, 4);  // file entry i is palette index i+1
        }
        pos += 12 + content + children * 0;  // children of non-MAIN chunks are laid out as following chunks
    }
    m.ok = have_size && have_voxels;
    return m;
}

std::vector<uint8_t> read_file(const std::string& path) {
    std::vector<uint8_t> buf;
    FILE* fp = std::fopen(path.c_str(), "rb");
    if (!fp) return buf;
    std::fseek(fp, 0, SEEK_END);
    long size = std::ftell(fp);
    std::fseek(fp, 0, SEEK_SET);
    if (size > 0) { buf.resize((size_t)size); if (std::fread(buf.data(), 1, (size_t)size, fp) != (size_t)size) buf.clear(); }
    std::fclose(fp);
    return buf;
}
}  // namespace

void init_grid_from_magika_voxel(Grid* grid, const std::string& path, glm::vec3 position, MaterialType type) {
    VoxModel m = parse_vox(read_file(path));
    if (!m.ok) std::cout << "Emtpy voxel model" << std::endl;
    grid->X = m.sx; grid->Y = m.sy; grid->Z = m.sz;
    grid->type = type;
    grid->num_grid_cells = m.sx * m.sy * m.sz;
    grid->cells.assign(grid->num_grid_cells, 0);
    grid->colors.assign(grid->num_grid_cells, glm::vec4(0, 0, 0, 0));
    grid->has_one_color_per_cell = true;
    int counter = 0;
    for (int x = 0; x < grid->X; x++)
        for (int y = 0; y < grid->Y; y++)
            for (int z = 0; z < grid->Z; z++) {
                uint8_t ci = m.voxels[x + (size_t)y * m.sx + (size_t)z * m.sx * m.sy];
                if (ci == 0) continue;
                int gi = x * grid->Y * grid->Z + y * grid->Z + z;
                grid->cells[gi] = ci;  // the palette index, 1..255 (src/VoxelLoader.cpp:79)
                grid->colors[gi] = glm::vec4(m.palette[ci][0] / 255.0f, m.palette[ci][1] / 255.0f, m.palette[ci][2] / 255.0f, m.palette[ci][3] / 255.0f);
                counter++;
            }
    grid->num_occupied_grid_cells = counter;
    grid->sparse_solid = true;
    grid->dynamic_solid = false;
    grid->position = position;
}

}  // namespace Lustrine
