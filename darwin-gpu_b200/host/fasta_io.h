// fasta_io.h -- FASTA reader with the reference's observable behaviour
// (fasta.cpp:19-98): names are the first header token split on any character
// that is not alphanumeric or '_'; sequence lines must be wrapped at exactly
// 70 columns (fasta.h:19), only the last line of a record may be shorter.
#pragma once
#include <string>
#include <vector>

namespace darwin {

struct FastaSet {
    std::vector<std::string> names;     // first header token (what the output lines print)
    std::vector<std::string> seqs;      // raw characters, case preserved
};

// Returns false (with a message in *err) on the conditions where the reference
// prints an error: unreadable file, sequence before the first header, wrong wrap.
bool read_fasta(const std::string &path, FastaSet *out, std::string *err);

// Reverse complement with the reference's alphabet (darwin.cpp:110-147):
// ACGT/acgt/N/n; any other character is an error.
bool reverse_complement(const std::string &seq, std::string *out, char *bad);

}  // namespace darwin
