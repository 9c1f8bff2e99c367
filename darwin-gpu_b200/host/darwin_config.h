// darwin_config.h -- params.cfg reader for the drop-in `darwin` CLI.
//
// Keeps the reference's configuration surface (params.cfg in the working
// directory; sections/keys of darwin.cpp:471-492; INI syntax and atof()
// semantics of ConfigFile.cpp:30-60 / Chameleon.cpp:88-90).
#pragma once
#include <map>
#include <string>

namespace darwin {

class IniFile {
public:
    // Missing file = empty map (the reference behaves the same way and then
    // throws on the first Value()).
    explicit IniFile(const std::string &path);
    // Throws std::runtime_error("<section>/<key> does not exist") when absent.
    double value(const std::string &section, const std::string &key) const;
    bool has(const std::string &section, const std::string &key) const;
private:
    std::map<std::string, std::string> kv_;
};

struct Params {
    // [GACT_scoring]
    int match = 1, mismatch = -1, gap_open = -1, gap_extend = -1;
    // [DSOFT_params]
    int seed_size = 14;
    unsigned bin_size = 64, window_size = 4;
    int threshold = 21, num_seeds = 800, seed_occurence_multiple = 32;
    int max_candidates = 1000000, num_nz_bins = 2500000;
    // [GACT_first_tile]   (first_tile_size is read and ignored, like darwin.cpp:487)
    int first_tile_size = 128, first_tile_score_threshold = 35;
    // [GACT_extend]
    int tile_size = 320, tile_overlap = 120;

    static Params from_file(const std::string &path);    // every key is required, as in the reference
};

}  // namespace darwin
