#include "darwin_config.h"

#include <cstdlib>
#include <fstream>
#include <stdexcept>

namespace darwin {

static std::string strip(const std::string &s)
{
    static const char *ws = " \t\r\n";
    const size_t b = s.find_first_not_of(ws);
    if (b == std::string::npos) return "";
    const size_t e = s.find_last_not_of(ws);
    return s.substr(b, e - b + 1);
}

IniFile::IniFile(const std::string &path)
{
    std::ifstream in(path.c_str());
    std::string line, section;
    while (std::getline(in, line)) {
        if (line.empty() || line[0] == '#' || line[0] == ';') continue;
        if (line[0] == '[') {
            const size_t close = line.find(']');
            section = strip(line.substr(1, close == std::string::npos ? std::string::npos : close - 1));
            continue;
        }
        const size_t eq = line.find('=');
        const std::string key = strip(line.substr(0, eq));
        const std::string val = (eq == std::string::npos) ? strip(line) : strip(line.substr(eq + 1));
        kv_[section + "/" + key] = val;
    }
}

bool IniFile::has(const std::string &section, const std::string &key) const
{
    return kv_.count(section + "/" + key) != 0;
}

double IniFile::value(const std::string &section, const std::string &key) const
{
    auto it = kv_.find(section + "/" + key);
    if (it == kv_.end()) throw std::runtime_error(section + "/" + key + " does not exist");
    return atof(it->second.c_str());
}

Params Params::from_file(const std::string &path)
{
    IniFile f(path);
    Params p;
    p.match = (int)f.value("GACT_scoring", "match");
    p.mismatch = (int)f.value("GACT_scoring", "mismatch");
    p.gap_open = (int)f.value("GACT_scoring", "gap_open");
    p.gap_extend = (int)f.value("GACT_scoring", "gap_extend");
    p.seed_size = (int)f.value("DSOFT_params", "seed_size");
    p.bin_size = (unsigned)f.value("DSOFT_params", "bin_size");
    p.window_size = (unsigned)f.value("DSOFT_params", "window_size");
    p.threshold = (int)f.value("DSOFT_params", "threshold");
    p.num_seeds = (int)f.value("DSOFT_params", "num_seeds");
    p.seed_occurence_multiple = (int)f.value("DSOFT_params", "seed_occurence_multiple");
    p.max_candidates = (int)f.value("DSOFT_params", "max_candidates");
    p.num_nz_bins = (int)f.value("DSOFT_params", "num_nz_bins");
    p.first_tile_size = (int)f.value("GACT_first_tile", "first_tile_size");
    p.first_tile_score_threshold = (int)f.value("GACT_first_tile", "first_tile_score_threshold");
    p.tile_size = (int)f.value("GACT_extend", "tile_size");
    p.tile_overlap = (int)f.value("GACT_extend", "tile_overlap");
    return p;
}

}  // namespace darwin
