// See MoFREAKUtilities.h.  Reference: src/MoFREAK/MoFREAKUtilities.cpp:374-498 (computeMoFREAKFromFile),
// :691-719 (writer), :1062-1134 (readMetadata), :1136-1212 (reader).
#include "MoFREAKUtilities.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <fstream>
#include <iostream>
#include <sstream>

namespace {

const char kMagic[8] = {'M', 'F', 'K', 'G', 'R', 'A', 'Y', '1'};

std::string baseName(const std::string& path) {          // boost::filesystem::path(filename).filename()
    const size_t p = path.find_last_of("/\\");
    return p == std::string::npos ? path : path.substr(p + 1);
}

std::vector<std::string> split(const std::string& s, char delim) {
    std::vector<std::string> elems;
    std::stringstream ss(s);
    std::string item;
    while (std::getline(ss, item, delim)) elems.push_back(item);
    return elems;
}

// operator<<(ostream&, float) at the default precision of 6 is "%g"
inline void appendFloat(std::string& out, float v) {
    char buf[48];
    const int n = snprintf(buf, sizeof buf, "%g", (double)v);
    out.append(buf, n);
}
inline void appendInt(std::string& out, long v) {
    char buf[24];
    const int n = snprintf(buf, sizeof buf, "%ld", v);
    out.append(buf, n);
}

}  // namespace

bool GrayVideo::load(const std::string& path, std::string* error) {
    std::ifstream f(path, std::ios::binary);
    char magic[8];
    int32_t hdr[3];
    if (!f || !f.read(magic, 8) || memcmp(magic, kMagic, 8) != 0 || !f.read((char*)hdr, 12) || hdr[0] <= 0 || hdr[1] <= 0 || hdr[2] < 0) {
        if (error) *error = "not a MFKGRAY1 gray video: " + path;
        return false;
    }
    width = hdr[0]; height = hdr[1]; frames = hdr[2];
    data.resize((size_t)width * height * frames);
    if (!data.empty() && !f.read((char*)data.data(), (std::streamsize)data.size())) {
        if (error) *error = "truncated gray video: " + path;
        return false;
    }
    return true;
}

bool GrayVideo::save(const std::string& path, const uint8_t* frames, int width, int height, int n, std::string* error) {
    std::ofstream f(path, std::ios::binary);
    const int32_t hdr[3] = {width, height, n};
    f.write(kMagic, 8);
    f.write((const char*)hdr, 12);
    f.write((const char*)frames, (std::streamsize)((size_t)width * height * n));
    if (!f) { if (error) *error = "cannot write " + path; return false; }
    return true;
}

void formatMoFREAKRow(std::string& out, float x, float y, int frame, float scale, float motion_x, float motion_y,
                      const unsigned* appearance, const unsigned* motion) {
    appendFloat(out, x); out += ' ';
    appendFloat(out, y); out += ' ';
    appendInt(out, frame); out += ' ';
    appendFloat(out, scale); out += ' ';
    appendFloat(out, motion_x); out += ' ';
    appendFloat(out, motion_y); out += ' ';
    for (int i = 0; i < APPEARANCE_BYTES; ++i) { appendInt(out, appearance[i]); out += ' '; }
    for (int i = 0; i < MOTION_BYTES; ++i) { appendInt(out, (int)motion[i]); out += ' '; }
    out += '\n';
}

MoFREAKUtilities::MoFREAKUtilities(int dset) : current_action(0), dataset(dset), device_(0), ctx_(nullptr), ctx_w_(0), ctx_h_(0), budget_scale_(1), rows_per_frame_hint_(1024) {}

MoFREAKUtilities::~MoFREAKUtilities() { mfk_destroy(ctx_); }

// the reference reports errors with cout + exit(1) (e.g. BagOfWordsRepresentation.cpp:147-151); so does this layer
void MoFREAKUtilities::die(const std::string& what) {
    std::cout << what << std::endl;
    exit(1);
}

void MoFREAKUtilities::ensureContext(int width, int height, int budget_scale) {
    if (budget_scale == 0) budget_scale = budget_scale_;
    if (ctx_ && ctx_w_ == width && ctx_h_ == height && budget_scale == budget_scale_) return;
    mfk_destroy(ctx_);
    ctx_ = nullptr;
    // a store window of 69 frames = 64 pairs per launch sequence plus the 5-frame look-back
    const int pairs = 64;
    // default budgets of the C ABI (pyramid pixels / 6 candidates, / 48 rows per pair), scaled up after an overflow; every
    // pyramid pixel being a candidate is the hard ceiling
    const long pyr = (long)(width * (double)height * 1.9);
    const int cand = budget_scale == 1 ? 0 : (int)std::min<long>(pyr, std::max<long>(4096, pyr / 6) * budget_scale);
    const int rows = budget_scale == 1 ? 0 : (int)std::min<long>(pyr, std::max<long>(1024, pyr / 48) * budget_scale);
    if (mfk_create_ex(&ctx_, device_, width, height, pairs + MFK_FRAME_GAP, pairs, cand, rows) != MFK_OK)
        die(std::string("mofreak_b200: ") + mfk_last_error(nullptr));
    ctx_w_ = width; ctx_h_ = height; budget_scale_ = budget_scale;
}

void MoFREAKUtilities::computeMoFREAKFromFile(std::string video_filename, std::string mofreak_filename, bool clear_features_after_computation) {
    GrayVideo v;
    std::string err;
    if (!v.load(video_filename, &err)) {
        std::cout << "Could not open file: " << video_filename << std::endl;      // MoFREAKUtilities.cpp:383-386
        die(err);
    }
    computeMoFREAKFromFrames(v.data.data(), v.width, v.height, v.frames, video_filename, mofreak_filename, clear_features_after_computation);
}

void MoFREAKUtilities::computeMoFREAKFromFrames(const uint8_t* gray, int width, int height, int n_frames, const std::string& name_for_metadata,
                                                const std::string& mofreak_filename, bool clear_features_after_computation) {
    ensureContext(width, height);
    // Rows AND their `.mofreak` text come back from the device (mfk_extract_clip_text): the reference's 22 ostream
    // tokens per row (:691-719) would otherwise cost far more host time than the whole extraction costs the GPU.
    size_t cap = (size_t)std::max(1, n_frames) * rows_per_frame_hint_, n = 0, n_text = 0;
    int rc;
    for (;;) {
        last_rows_.resize(cap);
        last_text_.resize(cap * 96);
        rc = mfk_extract_clip_text(ctx_, gray, (size_t)width, n_frames, &last_text_[0], last_text_.size(), &n_text, last_rows_.data(), cap, &n);
        if (rc == MFK_ERR_BUDGET && budget_scale_ < 64) {                   // a device-side budget overflowed (very busy frames):
            ensureContext(width, height, budget_scale_ * 4);                // the reference has no such limit, so grow and redo
            continue;
        }
        if (rc == MFK_ERR_CAPACITY) {
            // a buffer of OURS was too small; the device kept the accumulation, so read it again with the sizes it reported
            last_rows_.resize(std::max<size_t>(n, 1));
            last_text_.resize(std::max<size_t>(n_text, 1));
            const size_t want_rows = n;
            rc = mfk_rows_text(ctx_, &last_text_[0], last_text_.size(), &n_text);
            const int rc_rows = mfk_fetch_rows(ctx_, last_rows_.data(), last_rows_.size(), &n, nullptr, 0);
            if (rc == MFK_OK || rc == MFK_ERR_RANGE) { if (rc_rows != MFK_OK) rc = rc_rows; }
            if (rc == MFK_OK && n != want_rows) {                            // never silently write a short file
                rc = MFK_ERR_STATE;
                die("mofreak_b200: the device accumulation changed between the sizing call and the re-read");
            }
            break;
        }
        break;
    }
    const bool text_on_host = rc == MFK_ERR_RANGE;                          // a float "%g" prints in exponent notation: never for extracted rows
    if (rc != MFK_OK && !text_on_host) die(std::string("mofreak_b200: ") + mfk_last_error(ctx_));
    last_rows_.resize(n);
    last_text_.resize(text_on_host ? 0 : n_text);

    // gather metadata once per video (the reference re-parses the file name per keypoint, :468-469)
    int action = 0, person = 0, video_number = 0;
    readMetadata(name_for_metadata, action, video_number, person);
    const bool had_features = !features.empty();
    // the deque is only materialised when somebody can see it: the reference's callers pass clear = true (main.cpp:878, 918)
    if (!clear_features_after_computation || had_features || text_on_host) {
        for (const mfk_row& r : last_rows_) {
            MoFREAKFeature ftr(NUMBER_OF_BYTES_FOR_MOTION, NUMBER_OF_BYTES_FOR_APPEARANCE);
            ftr.frame_number = r.frame;
            ftr.scale = r.scale;
            ftr.x = r.x;
            ftr.y = r.y;
            for (int i = 0; i < NUMBER_OF_BYTES_FOR_APPEARANCE; ++i) ftr.appearance[i] = r.appearance[i];
            for (int i = 0; i < NUMBER_OF_BYTES_FOR_MOTION; ++i) ftr.motion[i] = r.motion[i];
            ftr.action = action; ftr.video_number = video_number; ftr.person = person;
            ftr.motion_x = 0; ftr.motion_y = 0;
            features.push_back(ftr);
        }
    }
    std::cout << "Writing this mofreak file: " << mofreak_filename << std::endl;
    if (had_features || text_on_host) writeMoFREAKFeaturesToFile(mofreak_filename);      // earlier features are written too, as in the reference
    else {
        std::ofstream f(mofreak_filename, std::ios::binary);
        f.write(last_text_.data(), (std::streamsize)last_text_.size());
        f.close();
    }
    if (clear_features_after_computation) features.clear();
}

void MoFREAKUtilities::writeMoFREAKFeaturesToFile(std::string output_file) {
    std::string text;
    text.reserve(features.size() * 96);
    for (auto it = features.begin(); it != features.end(); ++it)
        formatMoFREAKRow(text, it->x, it->y, it->frame_number, it->scale, it->motion_x, it->motion_y, it->appearance.data(), it->motion.data());
    std::ofstream f(output_file, std::ios::binary);
    f.write(text.data(), (std::streamsize)text.size());
    f.close();
}

void MoFREAKUtilities::readMetadata(std::string filename, int& action, int& video_number, int& person) {
    const std::string file_name_str = baseName(filename);
    // the KTH branch is disabled in the reference (`if (false)`, :1068); labels stay as they are there
    if (dataset == HMDB51) {
        video_number = 0;
        person = 0;
        action = current_action;
    } else if (dataset == UTI2) {
        const std::vector<std::string> parts = split(file_name_str, '_');
        if (parts.size() >= 3) {
            std::stringstream(parts[1]) >> person;
            std::stringstream(parts[2].substr(0, 1)) >> action;
            std::stringstream(parts[0]) >> video_number;
        }
    }
}

void MoFREAKUtilities::readMoFREAKFeatures(std::string filename, int num_to_sample) {
    std::deque<MoFREAKFeature> new_features;
    int action = 0, video_number = 0, person = 0;
    readMetadata(filename, action, video_number, person);

    std::ifstream stream;
    stream.open(filename);
    while (!stream.eof()) {
        MoFREAKFeature ftr(NUMBER_OF_BYTES_FOR_MOTION, NUMBER_OF_BYTES_FOR_APPEARANCE);
        stream >> ftr.x >> ftr.y >> ftr.frame_number >> ftr.scale >> ftr.motion_x >> ftr.motion_y;
        if (stream.eof() || stream.fail()) break;
        for (unsigned i = 0; i < (unsigned)NUMBER_OF_BYTES_FOR_APPEARANCE; ++i) { unsigned a = 0; stream >> a; ftr.appearance[i] = a; }
        for (unsigned i = 0; i < (unsigned)NUMBER_OF_BYTES_FOR_MOTION; ++i) { unsigned a = 0; stream >> a; ftr.motion[i] = a; }
        ftr.action = action; ftr.video_number = video_number; ftr.person = person;
        new_features.push_back(ftr);
    }
    stream.close();

    // the reference appends from the back of the freshly read list (:1185-1203), i.e. in reverse file order
    if (num_to_sample && ((int)new_features.size() > num_to_sample)) {
        std::random_shuffle(new_features.begin(), new_features.end());
        for (int i = 0; i < num_to_sample; ++i) { features.push_back(new_features.back()); new_features.pop_back(); }
    } else {
        while (!new_features.empty()) { features.push_back(new_features.back()); new_features.pop_back(); }
    }
}

std::deque<MoFREAKFeature> MoFREAKUtilities::getMoFREAKFeatures() { return features; }

void MoFREAKUtilities::setAllFeaturesToLabel(int label) {
    for (unsigned i = 0; i < features.size(); ++i) features[i].action = label;
}

void MoFREAKUtilities::clearFeatures() { features.clear(); }

// HMDB51 class folders in the order of the reference's HMDB_action enum (MoSIFTUtilities.h:15-21)
static const char* const kHmdb51[] = {
    "brush_hair", "cartwheel", "catch", "chew", "clap", "climb", "climb_stairs", "dive", "draw_sword", "dribble", "drink", "eat",
    "fall_floor", "fencing", "flic_flac", "golf", "handstand", "hit", "hug", "jump", "kick", "kick_ball", "kiss", "laugh", "pick",
    "pour", "pullup", "punch", "push", "pushup", "ride_bike", "ride_horse", "run", "shake_hands", "shoot_ball", "shoot_bow",
    "shoot_gun", "sit", "situp", "smile", "smoke", "somersault", "stand", "swing_baseball", "sword", "sword_exercise", "talk",
    "throw", "turn", "walk", "wave"};

void MoFREAKUtilities::setCurrentAction(std::string folder_name) {       // MoFREAKUtilities.cpp:782-1060
    if (dataset == HMDB51) {
        for (int i = 0; i < 51; i++)
            if (folder_name == kHmdb51[i]) { current_action = i; return; }
        current_action = 0;
        die("****Didn't find action");
    } else if (dataset == UCF101) {
        if (actions.find(folder_name) == actions.end()) { const int next = (int)actions.size(); actions[folder_name] = next; }
        current_action = actions[folder_name];
    }
}
