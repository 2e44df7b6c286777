// Command-line driver in the shape of the reference's main.cpp stages (main.cpp:854-924 computeMoFREAKFiles,
// :294-349 computeBOWRepresentation), with paths as arguments instead of compile-time globals:
//   mofreak_main detect <dataset> <out_dir> <video.y8>...                DETECT_MOFREAK
//   mofreak_main bow <dataset> <num_clusters> <num_groups> <svm_path> <file.mofreak>...
//   mofreak_main cluster <dataset> <num_clusters> <num_classes> <svm_path> <seed> <class>:<file.mofreak>...   cluster(), main.cpp:160-213
//   mofreak_main pipeline <dataset> <num_clusters> <num_classes> <num_groups> <work_dir> <seed> <class>:<video.y8>...
//       DETECTION_TO_CLASSIFICATION up to the SVM (main.cpp:944-962): computeMoFREAKFiles -> cluster -> computeBOWRepresentation;
//       writes <work_dir>/<stem>.mofreak, clusters.txt and <k>.train / <k>.test (classify() = libsvm stays on the host, out of scope)
// dataset: KTH | TRECVID | UTI2 | HMDB51 | UCF101
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <iostream>
#include <string>

#include <vector>

#include "BagOfWordsRepresentation.h"
#include "Clustering.h"
#include "MoFREAKUtilities.h"

static int datasetId(const std::string& s) {
    const char* names[] = {"KTH", "TRECVID", "HOLLYWOOD", "UTI1", "UTI2", "HMDB51", "UCF101"};
    for (int i = 0; i < 7; i++) if (s == names[i]) return i;
    std::cout << "unknown dataset " << s << std::endl;
    exit(1);
}

// cluster(), main.cpp:160-213: one class at a time (bounds memory), at most 50000 / n_files rows per file, randomClusters(true),
// one clusters.txt at the end.  files: "<class>:<path>"
static void clusterStage(int dset, int num_clusters, int num_classes, const std::string& svm_path, uint32_t seed,
                         const std::vector<std::string>& tagged) {
    std::vector<int> classes;
    for (int c = 1; c <= num_classes; c++) classes.push_back(c);
    Clustering clustering(16, num_clusters, 0, num_classes, classes, svm_path);
    clustering.setAppearanceDescriptor(8, true);
    clustering.setMotionDescriptor(8, true);
    clustering.setSeed(seed);
    for (int c = 1; c <= num_classes; c++) {
        MoFREAKUtilities mofreak(dset);
        std::vector<std::string> files;
        for (const std::string& a : tagged) {
            const size_t colon = a.find(':');
            if (colon != std::string::npos && atoi(a.substr(0, colon).c_str()) == c) files.push_back(a.substr(colon + 1));
        }
        if (files.empty()) continue;
        for (const std::string& f : files) mofreak.readMoFREAKFeatures(f, 50000 / (int)files.size());
        std::cout << "Building data." << std::endl;
        clustering.buildDataFromMoFREAK(mofreak.getMoFREAKFeatures(), false, false);
        clustering.randomClusters(true);
    }
    clustering.writeClusters(false);
}

static std::string stemOf(const std::string& path) {
    const size_t slash = path.find_last_of('/');
    std::string stem = slash == std::string::npos ? path : path.substr(slash + 1);
    const size_t dot = stem.find_last_of('.');
    return dot == std::string::npos ? stem : stem.substr(0, dot);
}

int main(int argc, char** argv) {
    if (argc < 4) { std::cout << "usage: see mofreak_main.cpp" << std::endl; return 2; }
    const auto start = std::chrono::steady_clock::now();
    const std::string mode = argv[1];
    const int dset = datasetId(argv[2]);
    if (mode == "detect") {
        MoFREAKUtilities mofreak(dset);
        const std::string out_dir = argv[3];
        for (int i = 4; i < argc; i++) {
            std::string video = argv[i];
            const size_t slash = video.find_last_of('/');
            std::string stem = slash == std::string::npos ? video : video.substr(slash + 1);
            const size_t dot = stem.find_last_of('.');
            if (dot != std::string::npos) stem = stem.substr(0, dot);
            std::cout << "FILE: " << video << std::endl;
            mofreak.computeMoFREAKFromFile(video, out_dir + "/" + stem + ".mofreak", true);
        }
    } else if (mode == "bow" && argc >= 7) {
        BagOfWordsRepresentation bow_rep(atoi(argv[3]), 16, argv[5], atoi(argv[4]), dset);
        bow_rep.intializeBOWMemory(argv[5]);
        for (int i = 6; i < argc; i++) bow_rep.convertFileToBOWFeature(argv[i]);
        bow_rep.writeBOWFeaturesToFiles();
    } else if (mode == "cluster" && argc >= 8) {
        std::vector<std::string> tagged(argv + 7, argv + argc);
        clusterStage(dset, atoi(argv[3]), atoi(argv[4]), argv[5], (uint32_t)strtoul(argv[6], nullptr, 10), tagged);
    } else if (mode == "pipeline" && argc >= 9) {
        const int num_clusters = atoi(argv[3]), num_classes = atoi(argv[4]), num_groups = atoi(argv[5]);
        const std::string work = argv[6];
        const uint32_t seed = (uint32_t)strtoul(argv[7], nullptr, 10);
        std::vector<std::string> tagged_mofreak, mofreak_files;
        {   // computeMoFREAKFiles, main.cpp:854-924
            MoFREAKUtilities mofreak(dset);
            for (int i = 8; i < argc; i++) {
                const std::string a = argv[i];
                const size_t colon = a.find(':');
                if (colon == std::string::npos) { std::cout << "expected <class>:<video>, got " << a << std::endl; return 2; }
                const std::string video = a.substr(colon + 1);
                const std::string out = work + "/" + stemOf(video) + ".mofreak";
                std::cout << "FILE: " << video << std::endl;
                mofreak.computeMoFREAKFromFile(video, out, true);
                tagged_mofreak.push_back(a.substr(0, colon + 1) + out);
                mofreak_files.push_back(out);
            }
        }
        clusterStage(dset, num_clusters, num_classes, work, seed, tagged_mofreak);
        {   // computeBOWRepresentation, main.cpp:294-349
            BagOfWordsRepresentation bow_rep(num_clusters, 16, work, num_groups, dset);
            bow_rep.intializeBOWMemory(work);
            for (const std::string& f : mofreak_files) bow_rep.convertFileToBOWFeature(f);
            bow_rep.writeBOWFeaturesToFiles();
        }
    } else {
        std::cout << "usage: see mofreak_main.cpp" << std::endl;
        return 2;
    }
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - start).count();
    std::cout << "Took this long: " << sec << " seconds" << std::endl;      // main.cpp:1011
    return 0;
}
