/*
 * pbr_headless -- the reference's main() (main.cpp:8-29) without Qt: load config.json, import a
 * model, render frames through PathTracer::generateImage, write the accumulated image.
 *
 *   pbr_headless [--config config.json] --model <dir/> <file.obj> [--frames N] [--out image.pfm]
 *                [--checkpoint acc.bin] [--resume acc.bin] [--deterministic] [--device D]
 *                [--ranks R [--shard spp|rows|stripes]] [--traversal -1|0|1] [--set key=value]...
 *
 * --ranks R: R processes, one per GPU (device D + rank), forked before anything touches CUDA; rank 0 creates the NCCL
 * id and hands it over through a file in /tmp; every rank loads the scene, PathTracer::setRanks does the rest (one
 * collective per frame inside libpbr_b200.so).  --frames counts frames per rank; rank 0 writes the image.
 */
#include <chrono>
#include <locale.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <string>
#include <vector>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "../../include/pbr_host.h"
#include "Cfg.h"
#include "Logger.h"
#include "qt/GLWidget.h"

int main( int argc, char** argv ) {
	setlocale( LC_ALL, "C" );

	std::string config, dir, file, out, checkpoint, resume;
	std::vector<std::string> sets;
	int frames = 1, device = -1, ranks = 1, shard = PathTracer::SHARD_SPP, traversal = -1;
	bool deterministic = false;

	for( int i = 1; i < argc; i++ ) {
		const std::string a = argv[i];
		if( a == "--config" && i + 1 < argc ) { config = argv[++i]; }
		else if( a == "--model" && i + 2 < argc ) { dir = argv[++i]; file = argv[++i]; }
		else if( a == "--frames" && i + 1 < argc ) { frames = atoi( argv[++i] ); }
		else if( a == "--out" && i + 1 < argc ) { out = argv[++i]; }
		else if( a == "--checkpoint" && i + 1 < argc ) { checkpoint = argv[++i]; }
		else if( a == "--resume" && i + 1 < argc ) { resume = argv[++i]; }
		else if( a == "--device" && i + 1 < argc ) { device = atoi( argv[++i] ); }
		else if( a == "--set" && i + 1 < argc ) { sets.push_back( argv[++i] ); }
		else if( a == "--deterministic" ) { deterministic = true; }
		else if( a == "--ranks" && i + 1 < argc ) { ranks = atoi( argv[++i] ); }
		else if( a == "--traversal" && i + 1 < argc ) { traversal = atoi( argv[++i] ); }
		else if( a == "--shard" && i + 1 < argc ) {
			const std::string v = argv[++i];
			shard = v == "rows" ? PathTracer::SHARD_ROWS : ( v == "stripes" ? PathTracer::SHARD_STRIPES : PathTracer::SHARD_SPP );
		}
		else {
			fprintf( stderr, "usage: %s [--config config.json] --model <dir/> <file.obj> [--frames N] [--out image.pfm]\n"
				"       [--checkpoint acc.bin] [--resume acc.bin] [--deterministic] [--device D]\n"
				"       [--ranks R [--shard spp|rows|stripes]] [--traversal -1|0|1] [--set key=value]...\n", argv[0] );
			return 2;
		}
	}
	if( file.empty() ) {
		fprintf( stderr, "no model given (--model <dir/> <file.obj>)\n" );
		return 2;
	}

	/* one process per GPU: fork before the first CUDA call */
	int rank = 0;
	std::vector<pid_t> children;
	char idPath[128];
	snprintf( idPath, sizeof( idPath ), "/tmp/pbr_nccl_id_%ld", (long) getpid() );
	if( ranks < 1 ) { ranks = 1; }
	for( int r = 1; r < ranks; r++ ) {
		const pid_t pid = fork();
		if( pid < 0 ) { perror( "fork" ); return 1; }
		if( pid == 0 ) { rank = r; children.clear(); break; }
		children.push_back( pid );
	}
	if( ranks > 1 ) { device = ( device < 0 ? 0 : device ) + rank; }

	try {
		if( !config.empty() ) { Cfg::get().loadConfigFile( config.c_str() ); }
		for( size_t i = 0; i < sets.size(); i++ ) {
			const size_t eq = sets[i].find( '=' );
			if( eq == std::string::npos ) { continue; }
			Cfg::get().value<std::string>( sets[i].substr( 0, eq ).c_str(), sets[i].substr( eq + 1 ) );
		}
	}
	catch( const std::exception& e ) {
		fprintf( stderr, "%s\n", e.what() );
		return 1;
	}

	CL::setDefaultDevice( device );
	int W = 0, H = 0;
	unsigned sampleCount = 0;
	std::vector<float> image;
	{
	/* (a scope of its own: the renderer -- and with it this rank's NCCL communicator -- is gone on EVERY rank before
	 *  rank 0 starts waiting for the others) */
	GLWidget widget;
	widget.getPathTracer()->setDeterministicSeeds( deterministic );
	widget.loadModel( dir, file );

	PathTracer* pt = widget.getPathTracer();
	pt->setTraversal( traversal );
	if( ranks > 1 ) {
		char id[128];
		const std::string tmp = std::string( idPath ) + ".tmp";
		if( rank == 0 ) {
			if( !CL::commUniqueId( id ) ) { fprintf( stderr, "NCCL is not available\n" ); return 1; }
			FILE* f = fopen( tmp.c_str(), "wb" );
			if( !f || fwrite( id, 1, 128, f ) != 128 || fclose( f ) != 0 || rename( tmp.c_str(), idPath ) != 0 ) {
				fprintf( stderr, "cannot write %s\n", idPath );
				return 1;
			}
		}
		else {
			bool have = false;
			for( int tries = 0; tries < 1200 && !have; tries++ ) {          /* up to two minutes: rank 0 is loading the scene too */
				FILE* f = fopen( idPath, "rb" );
				if( f ) { have = fread( id, 1, 128, f ) == 128; fclose( f ); }
				if( !have ) { usleep( 100000 ); }
			}
			if( !have ) { fprintf( stderr, "rank %d: no NCCL id from rank 0\n", rank ); return 1; }
		}
		if( !pt->setRanks( rank, ranks, id, shard ) ) { return 1; }
	}
	W = (int) pt->getWidth(); H = (int) pt->getHeight();
	image.assign( (size_t) W * H * 4, 0.0f );

	if( !resume.empty() ) {
		unsigned sc = 0;
		if( pbrh_read_checkpoint( resume.c_str(), image.data(), W, H, &sc ) != 0 ) {
			fprintf( stderr, "%s\n", pbrh_last_error() );
			return 1;
		}
		pt->writeImage( image.data(), sc );
	}

	const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
	if( frames > 1 ) { pt->renderFrames( (cl_uint) ( frames - 1 ) ); }
	if( frames > 0 ) { pt->generateImageInto( image.data(), NULL ); }
	const double sec = std::chrono::duration<double>( std::chrono::steady_clock::now() - t0 ).count();

	uint64_t stats[6];
	pt->getCL()->getStats( stats, false );
	fprintf( stderr, "%s%d frame(s) %dx%d in %.3f s: %.2f Msamples/s, %.2f Mrays/s, %u samples accumulated\n",
		ranks > 1 ? ( "rank " + std::to_string( rank ) + " of " + std::to_string( ranks ) + ": " ).c_str() : "",
		frames, W, H, sec, frames * (double) W * H * Cfg::get().value<int>( Cfg::RENDER_SAMPLES ) / sec * 1e-6,
		(double) ( stats[0] + stats[1] ) / sec * 1e-6, pt->getSampleCount() );
	sampleCount = pt->getSampleCount();
	pt->getCL()->finish();
	}
	if( rank != 0 ) { return 0; }
	int failed = 0;
	for( size_t i = 0; i < children.size(); i++ ) {
		int st = 0;
		if( waitpid( children[i], &st, 0 ) < 0 || !WIFEXITED( st ) || WEXITSTATUS( st ) != 0 ) { failed++; }
	}
	if( ranks > 1 ) { unlink( idPath ); }
	if( failed ) { fprintf( stderr, "%d rank(s) failed\n", failed ); return 1; }

	if( !out.empty() && pbrh_write_pfm( out.c_str(), image.data(), W, H ) != 0 ) {
		fprintf( stderr, "%s\n", pbrh_last_error() );
		return 1;
	}
	if( !checkpoint.empty() && pbrh_write_checkpoint( checkpoint.c_str(), image.data(), W, H, sampleCount ) != 0 ) {
		fprintf( stderr, "%s\n", pbrh_last_error() );
		return 1;
	}
	return 0;
}
