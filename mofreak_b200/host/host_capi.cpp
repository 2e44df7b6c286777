// extern "C" shim over the C++ host classes so that the Python test-suite can drive them (ctypes).
#include <string.h>

#include <string>

#include "BagOfWordsRepresentation.h"
#include "Clustering.h"
#include "MoFREAKUtilities.h"
#include "MotionInterchangePatterns.h"

extern "C" {

int mfkh_write_gray_video(const char* path, const uint8_t* frames, int w, int h, int n) {
    std::string err;
    return GrayVideo::save(path, frames, w, h, n, &err) ? 0 : -1;
}

// MoFREAKUtilities(dset).computeMoFREAKFromFile(video, mofreak, true); returns the row count
long mfkh_compute_mofreak_from_file(int dset, const char* video, const char* mofreak) {
    MoFREAKUtilities u(dset);
    u.computeMoFREAKFromFile(video, mofreak, false);
    return (long)u.getMoFREAKFeatures().size();
}

// the call the reference's drivers make (clear = true, main.cpp:878, 918): the file comes straight from the device-formatted
// text and no feature deque is built; rows_hint sizes the first host buffers (1 forces the re-read path)
long mfkh_compute_mofreak_clear(int dset, const char* video, const char* mofreak, long rows_hint) {
    MoFREAKUtilities u(dset);
    if (rows_hint > 0) u.setRowsPerFrameHint((size_t)rows_hint);
    u.computeMoFREAKFromFile(video, mofreak, true);
    return u.getMoFREAKFeatures().empty() ? (long)u.lastRows().size() : -1;
}

// reads a .mofreak file with the host reader and writes it back (round trip of reader + writer, in file order)
long mfkh_rewrite_mofreak(const char* in, const char* out) {
    MoFREAKUtilities u(MoFREAKUtilities::KTH);
    u.readMoFREAKFeatures(in);
    std::deque<MoFREAKFeature> f = u.getMoFREAKFeatures();
    MoFREAKUtilities v(MoFREAKUtilities::KTH);
    std::string text;
    for (auto it = f.rbegin(); it != f.rend(); ++it)          // the reader stores rows in reverse file order
        formatMoFREAKRow(text, it->x, it->y, it->frame_number, it->scale, it->motion_x, it->motion_y, it->appearance.data(), it->motion.data());
    FILE* fp = fopen(out, "wb");
    if (!fp) return -1;
    fwrite(text.data(), 1, text.size(), fp);
    fclose(fp);
    return (long)f.size();
}

// BagOfWordsRepresentation over files; writes SVM_PATH/<g>.train|.test.  Returns the codebook size.
int mfkh_bow_run(int num_clust, const char* svm_path, int num_groups, int dset, const char** files, int n_files) {
    BagOfWordsRepresentation bow(num_clust, 16, svm_path, num_groups, dset);
    bow.intializeBOWMemory(svm_path);
    for (int i = 0; i < n_files; i++) bow.convertFileToBOWFeature(files[i]);
    bow.writeBOWFeaturesToFiles();
    return (int)(bow.codebook().size() / 16);
}

int mfkh_bow_codebook(int num_clust, const char* svm_path, uint8_t* out, int cap_words) {
    BagOfWordsRepresentation bow(num_clust, 16, svm_path, 1, BagOfWordsRepresentation::KTH);
    const int K = (int)(bow.codebook().size() / 16);
    memcpy(out, bow.codebook().data(), (size_t)16 * (K < cap_words ? K : cap_words));
    return K;
}

int mfkh_bow_sliding(int num_clust, const char* svm_path, const char* mofreak, int alpha, int label, const char* out_path) {
    BagOfWordsRepresentation bow(num_clust, 16, svm_path, 1, BagOfWordsRepresentation::TRECVID);
    std::ofstream out(out_path);
    std::string f(mofreak);
    bow.computeSlidingBagOfWords(f, alpha, label, out);
    return 0;
}

int mfkh_mip(const uint8_t* cur, const uint8_t* prev, int w, int h, float scale, int x, int y, uint8_t* out8) {
    MotionInterchangePatterns mip(w, h);
    mip.setFrames(cur, prev);
    std::vector<unsigned int> d;
    mip.extractMotionByMotionInterchangePatterns(d, scale, x, y);
    for (int i = 0; i < 8; i++) out8[i] = (uint8_t)d[i];
    return 0;
}

// Clustering over .mofreak files grouped by class (the shape of cluster(), main.cpp:160-213): for each class its files are
// read (at most 50000 / n_files features per file), randomClusters(true) takes num_clust / num_classes of them, and
// clusters.txt is written once at the end.  seed 0 = time(NULL) as in the reference.
int mfkh_cluster(int num_clust, int num_classes, const char* svm_path, const char** files, const int* file_class, int n_files, unsigned seed) {
    std::vector<int> classes;
    for (int c = 1; c <= num_classes; c++) classes.push_back(c);
    Clustering clustering(16, num_clust, 0, num_classes, classes, svm_path);
    clustering.setAppearanceDescriptor(8, true);
    clustering.setMotionDescriptor(8, true);
    clustering.setSeed(seed);
    for (int c = 1; c <= num_classes; c++) {
        MoFREAKUtilities u(MoFREAKUtilities::KTH);
        int count = 0;
        for (int i = 0; i < n_files; i++) count += file_class[i] == c;
        if (!count) continue;
        for (int i = 0; i < n_files; i++) if (file_class[i] == c) u.readMoFREAKFeatures(files[i], 50000 / count);
        clustering.buildDataFromMoFREAK(u.getMoFREAKFeatures(), false, false);
        clustering.randomClusters(true);
    }
    clustering.writeClusters(false);
    return clustering.centersFilled();
}

}  // extern "C"
