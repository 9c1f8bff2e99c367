#include "fasta_io.h"

#include <cctype>
#include <fstream>

namespace darwin {

static const size_t WRAP = 70;

static std::string first_token(const std::string &header)
{
    std::string tok;
    for (size_t i = 1; i < header.size(); i++) {
        const unsigned char ch = (unsigned char)header[i];
        if (!isalpha(ch) && !isdigit(ch) && ch != '_') break;
        tok.push_back((char)ch);
    }
    return tok;
}

bool read_fasta(const std::string &path, FastaSet *out, std::string *err)
{
    std::ifstream in(path.c_str());
    if (!in.is_open()) { *err = "Error: Could not open FASTA file " + path + "."; return false; }
    std::string line, cur;
    bool have_record = false;
    size_t last_len = WRAP;
    while (std::getline(in, line)) {
        if (line.empty()) continue;
        if (line[0] == '>') {
            if (have_record) out->seqs.push_back(cur);
            cur.clear();
            out->names.push_back(first_token(line));
            have_record = true;
            last_len = WRAP;
            continue;
        }
        if (!have_record) { *err = "Error in file " + path + ": File begins with non-description line!"; return false; }
        if (line.size() > WRAP || (line.size() < WRAP && last_len != WRAP)) {
            *err = "Error in file " + path + ": FASTA sequence lines need to be wrapped to 70 characters!";
            return false;
        }
        cur += line;
        last_len = line.size();
    }
    if (have_record) out->seqs.push_back(cur);
    return true;
}

bool reverse_complement(const std::string &seq, std::string *out, char *bad)
{
    out->resize(seq.size());
    const size_t n = seq.size();
    for (size_t i = 0; i < n; i++) {
        char c = seq[n - 1 - i], r;
        switch (c) {
            case 'a': r = 't'; break; case 'A': r = 'T'; break;
            case 'c': r = 'g'; break; case 'C': r = 'G'; break;
            case 'g': r = 'c'; break; case 'G': r = 'C'; break;
            case 't': r = 'a'; break; case 'T': r = 'A'; break;
            case 'n': r = 'n'; break; case 'N': r = 'N'; break;
            default: if (bad) *bad = c; return false;
        }
        (*out)[i] = r;
    }
    return true;
}

}  // namespace darwin
