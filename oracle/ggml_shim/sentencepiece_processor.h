// Stand-in for <sentencepiece_processor.h> (oracle/_ref build, test infrastructure): the two calls the reference makes
// (Load, Encode: src/pocket_tts/conditioners/text.h:10-27, src/pocket_tts.cpp:308) over the repo's own C++ unigram encoder, whose ids
// are tested bit-exact against the upstream SentencePiece wheel (tests/test_text_frontend.py).
#pragma once
#include <string>
#include <vector>
#include "../../pocket-tts.cpp_b200/csrc/host/spm_unigram.hpp"
namespace sentencepiece {
namespace util { struct Status { bool good = true; bool ok() const { return good; } }; }
class SentencePieceProcessor {
public:
    util::Status Load(const std::string& path) { util::Status s; s.good = tok_.load(path); return s; }
    util::Status Encode(const std::string& text, std::vector<int>* ids) const { *ids = tok_.encode(text); return util::Status{}; }
    // only the unreachable offline helper split_into_best_sentences (conditioners/text.h:102-178) decodes; it has to compile, not run
    util::Status Decode(const std::vector<int>&, std::string*) const { fprintf(stderr, "sentencepiece stand-in: Decode is not supported\n"); abort(); }
private:
    mutable ptts_host::SpmUnigram tok_;
};
}  // namespace sentencepiece
